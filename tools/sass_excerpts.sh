#!/bin/bash
# SASS / PTX evidence of what the kernels compile to (no GPU needed): run from the repo root after build().
#   bash tools/sass_excerpts.sh > profiles/r2_sass_excerpts.txt
set -e
LIB=graphlily_b200/lib/libgraphlily_b200.so
SASS=$(mktemp); PTX=$(mktemp -d)
cuobjdump -sass $LIB > $SASS
fn() {  # fn <mangled-name regex>: SASS of the first matching function
    awk -v pat="$1" '/Function :/{f = ($0 ~ pat) && !done; if (f) done = 1} f' $SASS
}
mix() {  # instruction histogram of a function (mnemonic with modifiers)
    fn "$1" | grep -E '^\s+/\*[0-9a-f]{4}\*/' | sed -E 's/^\s+\/\*[0-9a-f]+\*\/\s+(@!?U?P[0-9T]+ )?//' | awk '{print $1}' | sort | uniq -c | sort -rn
}
echo "# $(cuobjdump -lelf $LIB | head -3 | tr '\n' ' ')"
echo "# nvcc $(nvcc --version | tail -2 | head -1)"
echo
echo "== spmv_lane_kernel<plus-times, unmasked, f32>: memory + shuffle instructions =="
mix 'spmv_lane_kernelILi0ELb0ELi0E' | grep -E 'LDG|STG|LDS|STS|SHFL|RED|ATOM|FMUL|FADD|FFMA|BAR|LDGSTS' 
echo "(two LDG.E.128 stream loads per group with .CONSTANT / no-allocate hints, 4-byte gathers with evict hints,"
echo " FMUL + FADD kept apart -- no FFMA on the value path: products and sums round separately like the host loop)"
echo
echo "== spmv_lane_tile_kernel<plus-times>: the optional TMA-tile variant =="
fn 'spmv_lane_tile_kernelILi0E' | grep -E 'UBLKCP|SYNCS|UTMA' | sed -E 's/\s+\/\* 0x[0-9a-f]+ \*\///' | head -8
echo
echo "== spmv_lane_bits_kernel<BITS=2, unmasked> (BFS or-and): memory instructions =="
mix 'spmv_lane_bits_kernelILi2ELb0E' | grep -E 'LDG|STG|LDS|STS|SHFL|POPC|LOP3|BAR' | head -14
echo
echo "== spmspv_kernel<plus-times, f32>: fire-and-forget reductions, barriers, conditional handle =="
mix 'spmspv_kernelILi0ELi0E' | grep -E 'RED|ATOM|LDG|STG|MEMBAR|SHFL|BAR|CALL' | head -24
echo
echo "== spmspv_kernel<min-plus, f32>: one integer RED.MIN per non-zero on order-preserving keys =="
mix 'spmspv_kernelILi2ELi0E' | grep -E 'RED|ATOM' | head -8
echo "-- destination register of the reductions (RZ = no value returned: the warp does not wait for the L2) --"
fn 'spmspv_kernelILi0ELi0E' | grep -E 'ATOMG' | sed -E 's/\s+\/\* 0x[0-9a-f]+ \*\///' | awk '{$1=""; print}' | sed -E 's/desc.*//' | sort | uniq -c | sort -rn | head -6
echo
echo "== xchg_push_signal_kernel: PTX (multimem.* never appears in SASS; the multicast address makes the STG a switch-replicated store) =="
nvcc -gencode arch=compute_100a,code=compute_100a -std=c++17 -Iinclude -Igraphlily_b200/csrc -ptx graphlily_b200/csrc/exchange.cu -o $PTX/exchange.ptx
grep -n 'multimem' $PTX/exchange.ptx
echo "-- SASS of the same kernel --"
fn 'xchg_push_signal_kernel' | grep -E 'STG|LDG|MEMBAR|ATOM|RED' | sed -E 's/\s+\/\* 0x[0-9a-f]+ \*\///'
echo
echo "== spmv_lane_kernel / spmv_fixup_kernel in-kernel pushes (GLB_XCHG_MC=fused|progressive): PTX =="
nvcc -gencode arch=compute_100a,code=compute_100a -std=c++17 -Iinclude -Igraphlily_b200/csrc -ptx graphlily_b200/csrc/spmv.cu -o $PTX/spmv.ptx
grep -c 'multimem.st' $PTX/spmv.ptx | sed 's/^/multimem.st occurrences in spmv.ptx: /'
grep -m3 'multimem' $PTX/spmv.ptx
echo
echo "== acquire at the head of gather_hot_kernel (ld.acquire.sys spin, trap on time-out) =="
fn 'gather_hot_kernel' | grep -E 'LDG.*STRONG|BPT|NANOSLEEP|CS2R|S2UR|LDG' | sed -E 's/\s+\/\* 0x[0-9a-f]+ \*\///' | head -10
echo
echo "== registers / shared memory (cuobjdump -res-usage) =="
cuobjdump -res-usage $LIB 2>/dev/null | grep -A1 -E 'spmv_lane_kernelILi0ELb0ELi0E|spmv_lane_bits_kernelILi2ELb0E|spmspv_kernelILi0ELi0E|spmv_fixup_kernelILi0ELi0E|xchg_push_signal' | grep -E 'Function|REG' | sed -E 's/Function ([^:]*):/\1/'
rm -rf $SASS $PTX
