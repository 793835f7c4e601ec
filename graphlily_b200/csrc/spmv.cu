// SpMV: y = A (+).(x) x over a row shard in lane-segment layout (overlay mode 1).
//
// Replaces kernel_spmv (/root/reference/graphlily/hw/kernel_spmv_impl.h:392-819) and the
// CPSR formatter that feeds it (graphlily/io/data_formatter.h:457-534).  Semantics are those of
// SpMVModule::compute_reference_results (graphlily/module/spmv_module.h:478-532).
//
// What bounds this kernel on B200 (tools/gather_bench.cu, profiles/r1_gather_microbench*.txt):
// a 4-byte gather that misses L1 costs one L1 miss request = one 32-byte L2 sector, and an SM
// sustains ~1.0 of those per clock whatever the path (ld / ld.cg / texture): 128 M random
// gathers take 464 us, 2.7x the 172 us the 1.12 GB of matrix stream needs from HBM.  Gathers
// that HIT L1 run 3x faster.  So the layout is built around keeping the gathers out of L2:
//
//   * hot columns: the formatter ranks the columns of the shard by reference count; the
//     tile_k most referenced ones are renumbered 0..tile_k-1 and every SpMV first packs their x
//     values into one contiguous `hot_x` array (a 200 KB gather) that the main kernel reads
//     with L1-allocating loads, so it stays L1-resident in every SM; all other columns keep
//     their id (stored as tile_k + column) and are read with ld.global.nc.L1::no_allocate, as
//     is the matrix stream, so they never evict the hot lines;
//   * lane-segment chunks: the nnz stream is cut into chunks of up to 8 groups of 128
//     non-zeros, one warp per chunk.  Inside a chunk lane l owns the 4n CONSECUTIVE non-zeros
//     [4n*l, 4n*(l+1)) and storage is transposed so that every load is still a coalesced
//     128-bit access (group g holds elements 4g..4g+3 of all 32 lanes: 512 B of column words
//     followed by 512 B of values).  A lane therefore reduces its run serially in registers --
//     no per-element shuffles or ballots -- and row boundaries come from one 32-bit flag word
//     per lane.  Row sums that close inside a lane go straight to a per-warp staging array in
//     shared memory indexed by row ordinal; one segmented shuffle scan per CHUNK (not per 128
//     non-zeros) stitches rows that cross lanes; then consecutive lanes write consecutive rows
//     (coalesced mask reads / y writes) with the mask and the fused eWiseAdd / dense-assign
//     epilogues applied in the same pass;
//   * a chunk never holds more than GLB_ROW_CAP row ends (the formatter cuts it early at a
//     row boundary and pads the last group; padding is flagged as a row of its own whose sum
//     is discarded), so the staging array is bounded;
//   * rows that cross or touch a chunk boundary leave per-chunk head / tail carries that the
//     fix-up kernel combines (deterministic -- no float atomics); it also writes empty rows.
//
// Algorithmic HBM bytes per launch: 8*nnz (cols+vals) + 4*nonempty_rows (nz_rows) + 4*ncols (x)
// + 4*rows (y) [+ 4*rows mask], i.e. the CSR figure of SURVEY.md section 8d; the flag words add
// 4 bytes per 32 non-zeros (1.6 %).
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <thread>
#include <vector>

#include "exchange.cuh"
#include "glb_internal.h"
#include "semiring.cuh"

namespace {

constexpr int kWarpsPerBlock = 8;
constexpr int kThreads = kWarpsPerBlock * 32;
constexpr unsigned kFull = 0xffffffffu;
#ifndef GLB_SPMV_WANT_CHUNKS
#define GLB_SPMV_WANT_CHUNKS 2960ull  // half a wave of warps: only shards below ~3 M non-zeros get smaller chunks
#endif
#ifndef GLB_SPMV_PREFETCH
#define GLB_SPMV_PREFETCH 2
#endif
#ifndef GLB_SPMV_MIN_BLOCKS
#define GLB_SPMV_MIN_BLOCKS 5
#endif
constexpr int kPrefetch = GLB_SPMV_PREFETCH;  // groups of the matrix stream in flight ahead of the one being reduced

struct SpmvParams {
    const uint32_t *__restrict__ stream;       // 256 words per group: 128 encoded columns, 128 fp32 values
    const uint32_t *__restrict__ flags;        // 32 per chunk: bit r of word l = "element r of lane l starts a row"
    const uint32_t *__restrict__ chunk_goff;   // n_chunks + 1: first group of every chunk
    const uint32_t *__restrict__ chunk_first;  // ordinal of the row open at the chunk start | GLB_FLAG if it starts there
    const uint32_t *__restrict__ nz_rows;
    const float *hot_x;   // x values of the hot columns, packed (or x itself when every column is hot)
    const float *x_cold;  // x - tile_k: cold column words index it directly
    const uint32_t *xbits;  // or-and only: bit w = (x value of stored column word w) != 0, see pack_bits_kernel
    // Shortcuts of the per-chunk chain of dependent loads (what a launch of a few waves is bound by: the BFS levels):
    const uint32_t *mbits;  // or-and with a mask: bit (row - mbits_base) = mask[row] != 0 at launch (pack_bits_kernel), an L1 hit
    uint32_t mbits_base;    //   instead of an L2 round trip per tested row
    int rows_identity;      // the shard has no empty row: the k-th non-empty row is row0 + k, nz_rows need not be read
    uint32_t row0;
    uint32_t uniform_groups, last_groups;  // every chunk but the last holds uniform_groups groups (0: read chunk_goff)
    const float *mask;    // may alias assign_inout
    float *y;
    float *y_peer[GLB_MAX_PEERS];  // the same vector on the other GPUs of a row-sharded run (peer-mapped memory)
    int n_peers;
    float *y_mc;  // or its multicast mapping: one store lands on every GPU (NVSwitch multicast)
    int mc_rows_in_main;  // 1: the main kernel's write-back stores every row to y_mc itself (GLB_XCHG_MC=fused)
    // Progressive push (multicast exchange, default): the CTA that completes a block of kPushCtas
    // consecutive CTAs copies the block's finished rows to y_mc in 16-byte multimem.st stores while
    // later blocks still compute; rows the fix-up kernel finishes are sent by that kernel, whose last
    // CTA then publishes the epoch.
    const uint32_t *push_bits;  // bit (row - push_row_base) = the main kernel writes this row
    const uint32_t *push_lo;    // row boundaries of the push blocks (n_push_blocks + 1)
    uint32_t *push_count;       // CTAs finished per push block (self-resetting)
    uint32_t push_row_base;
    uint32_t n_ctas;
    // Pusher CTAs (GLB_XCHG_MC=pusher): the main kernel only counts finished CTAs per push block;
    // xchg_pusher_kernel, resident beside it on SMs of its own, finishes the chunk-crossing / empty rows of every
    // completed block, sends the block and finally publishes (no fix-up launch).
    int push_count_only;
    uint32_t *pub_flags_mc;     // fix-up kernel: multicast mapping of the flag words (NULL: do not publish)
    uint32_t *pub_state;        // exchange state: [0] epoch, [1] CTA ticket
    int pub_rank;
    float *head_carry;
    float *tail_carry;
    uint32_t n_chunks;
    uint32_t chunk_begin, chunk_end;  // the chunks THIS launch of the main kernel covers (all, or one sub-block of a split step)
    uint32_t tile_k;
    float zero;
    int mask_type;
    int add_enable;
    float add_val;
    float *assign_inout;
    float assign_val;
    int assign_mask_type;
    // fix-up lists
    const glb_fixup_t *fix_short;
    const glb_fixup_t *fix_long;
    const uint32_t *empty_rows;
    uint32_t n_fix_short, n_fix_long, n_empty;
    uint32_t n_nz_rows;
};

// Everything that is read once (matrix stream, flags, row ids, mask) is kept out of L1 so it
// cannot evict the hot x lines.
__device__ __forceinline__ uint4 ld_stream_v4(const uint4 *p) {
    uint4 v;
    asm("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
        : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
        : "l"(p));
    return v;
}
__device__ __forceinline__ uint32_t ld_stream_u32(const uint32_t *p) {
    uint32_t v;
    asm("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}

__device__ __forceinline__ float ld_stream_f32(const float *p) {
    float v;
    asm volatile("ld.global.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p));  // mask may alias assign_inout: not .nc
    return v;
}

// Row write-back: fold `zero`, apply the mask (literal 0 compare / literal 0 write,
// spmv_module.h:513-532), then the optional fused eWiseAdd and dense assign.
template <int OP, bool IN_MAIN = true, int VT = GLB_VAL_F32>
__device__ __forceinline__ void finish_row(const SpmvParams &P, uint32_t row, float total) {
    float v = Semi<OP, VT>::with_zero(P.zero, total);
    if (VT == GLB_VAL_F32 && P.mbits) {  // (uniform) the mask as it was at launch, one bit per row
        const uint32_t b = row - P.mbits_base;
        const bool nz = (__ldg(P.mbits + (b >> 5)) >> (b & 31u)) & 1u;
        if (P.mask_type == GLB_MASK_WRITE_TO_ZERO ? nz : (P.mask_type == GLB_MASK_WRITE_TO_ONE && !nz)) v = Val<VT>::zero();
    } else if (P.mask_type == GLB_MASK_WRITE_TO_ZERO) {
        if (!Val<VT>::is_zero(ld_stream_f32(P.mask + row))) v = Val<VT>::zero();
    } else if (P.mask_type == GLB_MASK_WRITE_TO_ONE) {
        if (Val<VT>::is_zero(ld_stream_f32(P.mask + row))) v = Val<VT>::zero();
    }
    if (P.add_enable) v = Val<VT>::plus(v, P.add_val);
    P.y[row] = v;
    // fused exchange: the row also goes straight into every peer's copy over NVLink, so the
    // allgather of the next iteration's x rides inside the SpMV write-back
    if (P.y_mc && (!IN_MAIN || P.mc_rows_in_main)) {
        asm volatile("multimem.st.relaxed.sys.global.f32 [%0], %1;" ::"l"(P.y_mc + row), "f"(v) : "memory");
    } else {
#pragma unroll
        for (int p = 0; p < GLB_MAX_PEERS; ++p)
            if (p < P.n_peers) P.y_peer[p][row] = v;
    }
    if (P.assign_inout) {
        bool hit = (P.assign_mask_type == GLB_MASK_WRITE_TO_ONE) ? !Val<VT>::is_zero(v) : Val<VT>::is_zero(v);
        if (hit) P.assign_inout[row] = P.assign_val;
    }
}

// x gather: hot column words (< tile_k) index the packed hot vector through L1 (allocating),
// cold ones index x itself around L1.  Both addresses are formed, one load issues.
__device__ __forceinline__ float gather_x(const float *hot_x, const float *x_cold, uint32_t tile_k, uint32_t c) {
    float v;
    asm("{\n"
        " .reg .pred p;\n"
        " setp.lt.u32 p, %3, %4;\n"
        " @p ld.global.nc.L1::evict_last.f32 %0, [%1];\n"
        " @!p ld.global.nc.L1::no_allocate.f32 %0, [%2];\n"
        "}"
        : "=f"(v)
        : "l"(hot_x + c), "l"(x_cold + c), "r"(c), "r"(tile_k));
    return v;
}

// Shared-memory flavour of the gather: hot words index the CTA's copy of hot_x in shared memory.
__device__ __forceinline__ float gather_x_tile(uint32_t tile_base, const float *x_cold, uint32_t tile_k, uint32_t c) {
    float v;
    asm("{\n"
        " .reg .pred p;\n"
        " setp.lt.u32 p, %3, %4;\n"
        " @p ld.shared.f32 %0, [%1];\n"
        " @!p ld.global.nc.L1::no_allocate.f32 %0, [%2];\n"
        "}"
        : "=f"(v)
        : "r"(tile_base + c * 4u), "l"(x_cold + c), "r"(c), "r"(tile_k));
    return v;
}

// or-and flavour of the gather: only the truth value of x matters (spmv_module.h:497-500), so x is
// packed to one bit per stored column word (pack_bits_kernel) and the gather reads the 32-bit word
// holding it.  The bitmap is 1/32 of x: it stays L1-resident (all of it for graphs up to ~1.5 M
// columns), so these gathers are L1 hits instead of L2 sector requests.
__device__ __forceinline__ uint32_t gather_bit(const uint32_t *xbits, uint32_t c) {
    return (__ldg(xbits + (c >> 5)) >> (c & 31u)) & 1u;
}

__device__ __forceinline__ uint32_t row_of(const SpmvParams &P, uint32_t ord) {
    return P.rows_identity ? P.row0 + ord : ld_stream_u32(P.nz_rows + ord);
}
// mask[row] != 0, from the launch-time bitmap when there is one
__device__ __forceinline__ bool mask_nonzero(const SpmvParams &P, uint32_t row) {
    if (P.mbits) {
        const uint32_t b = row - P.mbits_base;
        return (__ldg(P.mbits + (b >> 5)) >> (b & 31u)) & 1u;
    }
    return ld_stream_f32(P.mask + row) != 0.0f;
}

// One chunk (up to 8 groups of 128 non-zeros) by one warp.  TILE: the hot vector is in shared
// memory at tile_base (persistent kernel), else it is read through L1.
// BITS (or-and only): 0 = fp32 x gathers, 1 = bitmap gathers + value stream (a != 0 is tested),
// 2 = bitmap gathers, pattern only (the formatter saw no stored zero: the value stream is not read).
// MASKED: the launch has a mask (compile-time so that unmasked launches carry none of the skip logic).
template <int OP, bool TILE, int BITS = 0, bool MASKED = true, int VT = GLB_VAL_F32>
__device__ __forceinline__ void process_chunk(const SpmvParams &P, const uint32_t chunk, const unsigned lane,
                                              float *const stage, const uint32_t tile_base) {
    uint32_t g0;
    int n;  // 1 .. GLB_MAX_GROUPS, warp-uniform
    if (P.uniform_groups) {
        g0 = chunk * P.uniform_groups;
        n = int(chunk + 1 == P.n_chunks ? P.last_groups : P.uniform_groups);
    } else {
        g0 = ld_stream_u32(P.chunk_goff + chunk);
        n = int(ld_stream_u32(P.chunk_goff + chunk + 1) - g0);
    }
    const uint4 *gp = reinterpret_cast<const uint4 *>(P.stream) + size_t(g0) * 64 + lane;
    uint4 cq[kPrefetch + 1], aq[kPrefetch + 1];
#pragma unroll
    for (int g = 0; g < kPrefetch; ++g) {
        if (g < n) {
            cq[g] = ld_stream_v4(gp + g * 64);
            if (BITS != 2) aq[g] = ld_stream_v4(gp + g * 64 + 32);
        }
    }
    const uint32_t fw = ld_stream_u32(P.flags + size_t(chunk) * 32 + lane);
    const uint32_t cf = ld_stream_u32(P.chunk_first + chunk);

    // ordinal of this lane's first row end = number of row ends in the lanes below
    const uint32_t cnt = __popc(fw);
    uint32_t incl = cnt;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t t = __shfl_up_sync(kFull, incl, d);
        if (int(lane) >= d) incl += t;
    }
    const uint32_t excl = incl - cnt;
    const uint32_t total = __shfl_sync(kFull, incl, 31);

    // Rows the mask discards need no gathers (the pull BFS case: from the third level on almost
    // every vertex is already visited).  The lane's run covers rows excl .. excl + cnt of the chunk;
    // it is "live" unless every one of them is masked out (lanes with many short rows do not check).
    // The mask vector may be the fused assign's inout, but a row is only ever assigned by the warp
    // (or the fix-up launch) that finishes it, after these reads.
    bool lane_live = true;
    if (MASKED && P.mask_type != GLB_MASK_NONE && cnt <= 3u) {
        lane_live = false;
        const uint32_t ord_base = (cf & ~GLB_FLAG) + excl;
        for (uint32_t j = 0; j <= cnt; ++j) {
            const uint32_t ord = ord_base + j;
            if (ord >= P.n_nz_rows) continue;  // the padding pseudo-row of the last chunk
            if (VT == GLB_VAL_F32) {
                const bool nz = mask_nonzero(P, row_of(P, ord));
                lane_live |= (P.mask_type == GLB_MASK_WRITE_TO_ZERO) ? !nz : nz;
            } else {
                const float mv = ld_stream_f32(P.mask + row_of(P, ord));
                lane_live |= (P.mask_type == GLB_MASK_WRITE_TO_ZERO) ? Val<VT>::is_zero(mv) : !Val<VT>::is_zero(mv);
            }
        }
    }
    const bool chunk_live = !MASKED || __any_sync(kFull, lane_live);

    // the lane's run: serial reduction in registers; every flagged element closes the open row
    float *sp = stage + excl;
    float acc = Semi<OP, VT>::ident();
    uint32_t run_bits = 0;  // BITS only
#pragma unroll
    for (int g = 0; g < GLB_MAX_GROUPS; ++g) {
        if (g < n) {
            if (chunk_live && g + kPrefetch < GLB_MAX_GROUPS && g + kPrefetch < n) {
                cq[(g + kPrefetch) % (kPrefetch + 1)] = ld_stream_v4(gp + (g + kPrefetch) * 64);
                if (BITS != 2) aq[(g + kPrefetch) % (kPrefetch + 1)] = ld_stream_v4(gp + (g + kPrefetch) * 64 + 32);
            }
            const uint4 c4 = cq[g % (kPrefetch + 1)];
            const uint4 a4 = (BITS != 2) ? aq[g % (kPrefetch + 1)] : make_uint4(0u, 0u, 0u, 0u);
            const uint32_t c[4] = {c4.x, c4.y, c4.z, c4.w};
            const uint32_t a[4] = {a4.x, a4.y, a4.z, a4.w};
            if (BITS) {
                // or-and over a bitmap x: a product is ONE bit.  The lane collects the bits of its run in a word (element r at
                // bit r) and cuts it into rows afterwards with the flag word -- ~6 instructions per non-zero instead of the
                // ~30 of the generic reduce-and-test-the-flag loop, which left this kernel issue-bound (56 % issue slots busy).
                if (!MASKED || lane_live) {
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        uint32_t bit = gather_bit(P.xbits, c[e]);
                        if (BITS == 1) bit &= (__uint_as_float(a[e]) != 0.0f) ? 1u : 0u;  // a stored 0.0 contributes nothing
                        run_bits |= bit << (4 * g + e);
                    }
                }
                continue;
            }
            float xv[4] = {0.0f, 0.0f, 0.0f, 0.0f};
            if (!MASKED || lane_live) {
#pragma unroll
                for (int e = 0; e < 4; ++e)
                    xv[e] = TILE ? gather_x_tile(tile_base, P.x_cold, P.tile_k, c[e]) : gather_x(P.hot_x, P.x_cold, P.tile_k, c[e]);
            }
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                float prod = Semi<OP, VT>::mul(__uint_as_float(a[e]), xv[e]);
                if (MASKED && !lane_live) prod = Semi<OP, VT>::ident();  // every row of this lane is masked out
                if (fw & (1u << (4 * g + e))) {
                    *sp++ = acc;
                    acc = Semi<OP, VT>::ident();
                }
                acc = Semi<OP, VT>::add(acc, prod);
            }
        }
    }
    if (BITS) {
        // rows of the run: the k-th flag (element p starts a new row) closes the segment [lo, p); the tail after the last
        // flag stays open in `acc`.  OR over a segment = "any bit of run_bits inside it".
        uint32_t f = fw, lo_mask = 0xffffffffu;  // lo_mask: bits at and above the start of the open segment
        while (f) {
            const uint32_t p = __ffs(int(f)) - 1;
            f &= f - 1;
            const uint32_t below_p = (1u << p) - 1u;
            *sp++ = (run_bits & below_p & lo_mask) ? 1.0f : 0.0f;
            lo_mask = ~below_p;
        }
        acc = (run_bits & lo_mask) ? 1.0f : 0.0f;
    }

    // stitch rows that cross lanes: segmented inclusive scan of the lane tails (a lane holding a
    // row start begins a new segment with what follows its last flag)
    const bool hasf = fw != 0;
    const unsigned hb = __ballot_sync(kFull, hasf);
    const unsigned le_mask = 0xffffffffu >> (31u - lane);
    int start = 31 - __clz(int(hb & le_mask));
    if (start < 0) start = 0;
    float v = acc;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const float tv = __shfl_up_sync(kFull, v, d);
        if (int(lane) - d >= start) v = Semi<OP, VT>::add(tv, v);
    }
    float carry_in = __shfl_up_sync(kFull, v, 1);
    if (lane == 0) carry_in = Semi<OP, VT>::ident();
    if (hasf) stage[excl] = Semi<OP, VT>::add(carry_in, stage[excl]);  // the lane's first row end began in lower lanes
    if (lane == 31) P.tail_carry[chunk] = v;
    __syncwarp();

    // write back: the k-th row end of the chunk closes row ordinal ord0 + k
    const uint32_t ord0 = cf & ~GLB_FLAG;
    const bool fresh = (cf & GLB_FLAG) != 0;
    for (uint32_t k = lane; k < total; k += 32) {
        const float val = stage[k];
        if (k == 0 && !fresh) {
            P.head_carry[chunk] = val;  // row began in an earlier chunk
        } else {
            finish_row<OP, true, VT>(P, row_of(P, ord0 + k), val);
        }
    }
}

constexpr uint32_t kPushCtas = 64;  // CTAs per push block: 512 chunks, ~64 KB of y on a 32-nnz/row matrix

// Tail of the main kernels in a row-sharded run over a multicast exchange: count this CTA as done;
// the CTA that completes its push block sends the block's rows (those the main kernel writes: the
// push_bits filter leaves chunk-crossing and empty rows to the fix-up kernel) to every rank.
__device__ __forceinline__ void push_block_when_complete(const SpmvParams &P) {
    __shared__ uint32_t s_last;
    const uint32_t blk = blockIdx.x / kPushCtas;
    __syncthreads();  // every warp of the CTA has written its rows
    if (threadIdx.x == 0) {
        const uint32_t first = blk * kPushCtas;
        const uint32_t in_block = P.n_ctas - first < kPushCtas ? P.n_ctas - first : kPushCtas;
        if (P.push_count_only) {
            // release at GPU scope, cumulative over the barrier above; no value needed.  (__threadfence + atomicAdd here --
            // a MEMBAR.SC in every CTA's exit -- stretched the main kernel from 186 to 205 us on half of C2)
            asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(P.push_count + blk) : "memory");
        } else {
            __threadfence();
            s_last = atomicAdd(P.push_count + blk, 1u) == in_block - 1;
        }
    }
    if (P.push_count_only) return;  // uniform: the pusher CTAs watch the counter
    __syncthreads();
    if (!s_last) return;
    __threadfence();  // the other CTAs' rows (counted above) are visible from here on
    const uint32_t lo = P.push_lo[blk], hi = P.push_lo[blk + 1];
    for (uint32_t r0 = (lo & ~3u) + 4u * threadIdx.x; r0 < hi; r0 += 4u * kThreads) {
        const uint32_t rel = r0 - P.push_row_base;
        uint32_t bits = (__ldg(P.push_bits + (rel >> 5)) >> (rel & 31u)) & 0xfu;
#pragma unroll
        for (int i = 0; i < 4; ++i)
            if (r0 + i < lo || r0 + i >= hi) bits &= ~(1u << i);
        if (bits == 0xfu) {
            const float4 v = __ldcg(reinterpret_cast<const float4 *>(P.y + r0));
            asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(P.y_mc + r0), "f"(v.x), "f"(v.y),
                         "f"(v.z), "f"(v.w)
                         : "memory");
        } else {
#pragma unroll
            for (int i = 0; i < 4; ++i)
                if (bits & (1u << i)) {
                    const float v = __ldcg(P.y + r0 + i);
                    asm volatile("multimem.st.relaxed.sys.global.f32 [%0], %1;" ::"l"(P.y_mc + r0 + i), "f"(v) : "memory");
                }
        }
    }
    if (threadIdx.x == 0) P.push_count[blk] = 0;  // ready for the next launch
}

// Variant L1: one chunk per warp, hot x lines kept in L1 by the cache hints.
template <int OP, bool MASKED, int VT = GLB_VAL_F32>
__global__ void __launch_bounds__(kThreads, GLB_SPMV_MIN_BLOCKS) spmv_lane_kernel(const SpmvParams P) {
    __shared__ float stage_all[kWarpsPerBlock][GLB_ROW_CAP];
    const unsigned lane = threadIdx.x & 31u;
    const unsigned wib = threadIdx.x >> 5;
    const uint32_t chunk = P.chunk_begin + blockIdx.x * kWarpsPerBlock + wib;
    if (chunk < P.chunk_end) process_chunk<OP, false, 0, MASKED, VT>(P, chunk, lane, stage_all[wib], 0u);  // warp-uniform
    if (P.push_bits) push_block_when_complete(P);
}

// Variant BITS (or-and): as above with x packed to a bitmap by pack_bits_kernel.
template <int BITS, bool MASKED>
__global__ void __launch_bounds__(kThreads, GLB_SPMV_MIN_BLOCKS) spmv_lane_bits_kernel(const SpmvParams P) {
    __shared__ float stage_all[kWarpsPerBlock][GLB_ROW_CAP];
    const unsigned lane = threadIdx.x & 31u;
    const unsigned wib = threadIdx.x >> 5;
    const uint32_t chunk = P.chunk_begin + blockIdx.x * kWarpsPerBlock + wib;
    if (chunk < P.chunk_end) process_chunk<GLB_OP_LOGICAL_AND_OR, false, BITS, MASKED>(P, chunk, lane, stage_all[wib], 0u);
    if (P.push_bits) push_block_when_complete(P);
}

// xbits: bit w = truth value of the x entry that stored column word w refers to.  Words below
// tile_k are hot ranks (x[hot_cols[w]], or x[w] under the identity numbering), the others are
// tile_k + column.  One warp packs 32 consecutive words with a ballot.
// `wait`: in a row-sharded run x is complete only once every rank's slice of the previous step has
// landed (glb_xchg_wait_head); x is then read around L1 (ld.cg) -- peers wrote it.
__global__ void __launch_bounds__(kThreads) pack_bits_kernel(const float *x, const uint32_t *__restrict__ hot_cols,
                                                           uint32_t *__restrict__ xbits, uint32_t tile_k, uint32_t n_hot,
                                                           uint32_t num_cols, uint32_t n_words32, const GlbXchgWait wait,
                                                           const float *mask, uint32_t *__restrict__ mbits, uint32_t mbits_base,
                                                           uint32_t row_begin, uint32_t row_end, uint32_t n_mwords) {
    glb_xchg_wait_head(wait);
    const uint32_t w = blockIdx.x * kThreads + threadIdx.x;  // grid covers (n_words32 + n_mwords) * 32 exactly
    if (w >= n_words32 * 32u) {  // warp-uniform: the mask of a masked launch, one bit per row of the shard
        const uint32_t row = mbits_base + (w - n_words32 * 32u);
        const bool nz = ((w >> 5) - n_words32) < n_mwords && row >= row_begin && row < row_end && mask[row] != 0.0f;
        const unsigned mb = __ballot_sync(kFull, nz);
        if ((threadIdx.x & 31u) == 0 && ((w >> 5) - n_words32) < n_mwords) mbits[(w >> 5) - n_words32] = mb;
        return;
    }
    bool t = false;
    if (w < tile_k) {
        const uint32_t c = n_hot ? __ldg(hot_cols + w) : w;
        t = (c < num_cols) && __ldcg(x + c) != 0.0f;
    } else if (w - tile_k < num_cols) {
        t = __ldcg(x + (w - tile_k)) != 0.0f;
    }
    const unsigned b = __ballot_sync(kFull, t);
    if ((threadIdx.x & 31u) == 0 && (w >> 5) < n_words32) xbits[w >> 5] = b;
}

// ---- TMA (bulk async copy) + mbarrier helpers -------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return uint32_t(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t phase) {
    uint32_t done;
    do {
        asm volatile(
            "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(phase)
            : "memory");
    } while (!done);
}

// Variant TILE: persistent CTAs, one per SM.  The CTA pulls the packed hot vector into shared
// memory with TMA bulk copies once (192 KB from L2), then its warps walk the chunk list round
// robin; hot gathers are bank-parallel LDS with no eviction, cold ones go around L1 as before.
template <int OP>
__global__ void __launch_bounds__(1024, 1) spmv_lane_tile_kernel(const SpmvParams P) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float *tile = reinterpret_cast<float *>(smem_raw);
    const unsigned n_warps = blockDim.x >> 5;
    float *stage = tile + ((P.tile_k + 3u) & ~3u);
    uint64_t *bar = reinterpret_cast<uint64_t *>(stage + n_warps * GLB_ROW_CAP);
    const unsigned lane = threadIdx.x & 31u;
    const unsigned wib = threadIdx.x >> 5;

    const uint32_t bytes = P.tile_k * 4u;
    const bool bulk_ok = (bytes % 16u == 0) && ((reinterpret_cast<uintptr_t>(P.hot_x) & 15u) == 0);
    if (bulk_ok) {
        if (threadIdx.x == 0) mbar_init(bar, 1);
        __syncthreads();
        if (threadIdx.x == 0) {
            mbar_expect_tx(bar, bytes);
            constexpr uint32_t kPiece = 32768;
            for (uint32_t off = 0; off < bytes; off += kPiece) {
                const uint32_t nb = (bytes - off < kPiece) ? (bytes - off) : kPiece;
                bulk_g2s(reinterpret_cast<unsigned char *>(tile) + off,
                         reinterpret_cast<const unsigned char *>(P.hot_x) + off, nb, bar);
            }
        }
        mbar_wait(bar, 0);
    } else {  // unaligned caller vector under the identity numbering: plain loads
        for (uint32_t i = threadIdx.x; i < P.tile_k; i += blockDim.x) tile[i] = P.hot_x[i];
        __syncthreads();
    }

    const uint32_t tile_base = smem_u32(tile);
    const uint32_t total_warps = gridDim.x * n_warps;
    for (uint32_t chunk = blockIdx.x * n_warps + wib; chunk < P.n_chunks; chunk += total_warps) {
        process_chunk<OP, true>(P, chunk, lane, stage + wib * GLB_ROW_CAP, tile_base);
        __syncwarp();  // the staging array is reused by the next chunk
    }
}

// hot_x[i] = x[hot_cols[i]]: the x values of the most referenced columns, packed.
__global__ void __launch_bounds__(kThreads) gather_hot_kernel(const float *x, const uint32_t *__restrict__ hot_cols,
                                                            float *__restrict__ hot_x, uint32_t n, const GlbXchgWait wait) {
    glb_xchg_wait_head(wait);  // row-sharded run: the peers' slices of x must have landed (see pack_bits_kernel)
    const uint32_t i = blockIdx.x * kThreads + threadIdx.x;
    if (i < n) hot_x[i] = __ldcg(x + __ldg(hot_cols + i));
}

// Rows touching a chunk boundary: total = tail[c_begin .. c_last] (+) head[c_end].  In a row-sharded
// run over a multicast exchange these rows (and the empty ones) go to every rank from here, and the
// CTA that finishes last publishes the epoch of the step (the main kernel's pushes completed with
// that kernel, before this one started).
template <int OP, int VT = GLB_VAL_F32>
__global__ void __launch_bounds__(kThreads) spmv_fixup_kernel(const SpmvParams P, uint32_t nb_short, uint32_t nb_long) {
    const unsigned lane = threadIdx.x & 31u;
    if (blockIdx.x < nb_short) {
        const uint32_t i = blockIdx.x * kThreads + threadIdx.x;
        if (i < P.n_fix_short) {
            const glb_fixup_t e = P.fix_short[i];
            const uint32_t c_end = e.c_end & ~GLB_FLAG;
            const bool has_head = (e.c_end & GLB_FLAG) != 0;
            const uint32_t c_stop = has_head ? c_end : c_end + 1;  // exclusive end of the tail range
            float t = Semi<OP, VT>::ident();
            for (uint32_t c = e.c_begin; c < c_stop; ++c) t = Semi<OP, VT>::add(t, P.tail_carry[c]);
            if (has_head) t = Semi<OP, VT>::add(t, P.head_carry[c_end]);
            finish_row<OP, false, VT>(P, e.row, t);
        }
    } else if (blockIdx.x < nb_short + nb_long) {
        const uint32_t i = (blockIdx.x - nb_short) * kWarpsPerBlock + (threadIdx.x >> 5);
        if (i < P.n_fix_long) {  // warp-uniform
            const glb_fixup_t e = P.fix_long[i];
            const uint32_t c_end = e.c_end & ~GLB_FLAG;
            const bool has_head = (e.c_end & GLB_FLAG) != 0;
            const uint32_t c_stop = has_head ? c_end : c_end + 1;
            float t = Semi<OP, VT>::ident();
            for (uint32_t c = e.c_begin + lane; c < c_stop; c += 32) t = Semi<OP, VT>::add(t, P.tail_carry[c]);
#pragma unroll
            for (int d = 16; d >= 1; d >>= 1) t = Semi<OP, VT>::add(t, __shfl_xor_sync(kFull, t, d));
            if (lane == 0) {
                if (has_head) t = Semi<OP, VT>::add(t, P.head_carry[c_end]);
                finish_row<OP, false, VT>(P, e.row, t);
            }
        }
    } else {
        const uint32_t i = (blockIdx.x - nb_short - nb_long) * kThreads + threadIdx.x;
        if (i < P.n_empty) finish_row<OP, false, VT>(P, P.empty_rows[i], Semi<OP, VT>::ident());
    }
    if (P.pub_flags_mc) {  // uniform
        __threadfence_system();
        __syncthreads();
        if (threadIdx.x == 0 && atomicAdd(P.pub_state + 1, 1u) == gridDim.x - 1) {
            P.pub_state[1] = 0;
            __threadfence_system();
            const uint32_t epoch = P.pub_state[0] + 1;
            P.pub_state[0] = epoch;
            asm volatile("multimem.st.release.sys.global.u32 [%0], %1;" ::"l"(P.pub_flags_mc + P.pub_rank), "r"(epoch) : "memory");
        }
    }
}

// Pusher CTAs of a row-sharded step over a multicast exchange (GLB_XCHG_MC=pusher).  A handful of CTAs, each holding
// an SM to itself (its shared-memory request leaves no room for a compute CTA), resident from the start of the step:
// remote stores issued from SMs that ALSO run compute CTAs hold up those SMs' L1 request port -- the unit that bounds
// the SpMV -- so the stores come from SMs that do nothing else.  CTA g watches push blocks g, g + G, ...: when all
// kPushCtas main-kernel CTAs of a block have counted themselves done it
//   1. finishes the block's chunk-crossing rows whose chunks all lie inside the block, and its empty rows (what the
//      fix-up kernel does otherwise: there is NO fix-up launch in this mode),
//   2. copies the block's rows to the multicast address in 16-byte multimem.st stores (the switch replicates them
//      into every rank's x) while the later blocks still compute.
// Once every block is out, the few rows that cross a block boundary (and the hub rows spanning > 32 chunks) are
// finished and sent as scalars, every CTA fences its stores system-wide and the last one publishes the step's epoch.
// Measured (2 GPUs, C2): a scalar multicast store costs an SM ~4 ns, so rows must travel in 16-byte pieces -- leaving
// all chunk-crossing rows (1 in 43) to a later scalar pass cost 60 us per step.
struct PusherParams {
    float *y_mc;
    const uint32_t *push_bits, *push_lo;  // bit clear = the row is sent late (crosses a block boundary / hub row)
    uint32_t *push_count, *pusher_done;
    const glb_fixup_t *fix_in, *fix_def;  // chunk-crossing rows inside one block (grouped by block) / across blocks
    const uint32_t *blk_fs, *blk_em;      // per block: range of fix_in / of empty_rows
    uint32_t n_fix_def;
    uint32_t push_row_base, n_blocks, n_ctas;
    uint32_t *mc_flags, *state;
    int rank;
    unsigned long long *trace;  // GLB_XCHG_TRACE: [0] start [1] all blocks out [2] late rows fenced [3] published [4+2b] block b ready [5+2b] sent
};
constexpr uint32_t kPusherThreads = 1024;

__device__ __forceinline__ unsigned long long glb_now_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

__device__ __forceinline__ void pusher_spin(const uint32_t *counter, uint32_t want, uint32_t *err) {
    const long long t0 = clock64();
    uint32_t seen;
    for (;;) {
        asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(counter) : "memory");
        if (seen >= want) break;
        __nanosleep(200);
        if (clock64() - t0 > 20000000000ll) {  // ~10 s: the SpMV kernels of this step never ran (see glb_xchg_spin)
            *err = 1;
            __threadfence_system();
            asm volatile("trap;");
        }
    }
}

// One chunk-crossing row, one thread (carries read around L1: the main kernel wrote them during THIS kernel's life).
template <int OP>
__device__ __forceinline__ void pusher_fix_short(const SpmvParams &P, const glb_fixup_t e) {
    const uint32_t c_end = e.c_end & ~GLB_FLAG;
    const bool has_head = (e.c_end & GLB_FLAG) != 0;
    const uint32_t c_stop = has_head ? c_end : c_end + 1;
    float t = Semi<OP>::ident();
    for (uint32_t c = e.c_begin; c < c_stop; ++c) t = Semi<OP>::add(t, __ldcg(P.tail_carry + c));
    if (has_head) t = Semi<OP>::add(t, __ldcg(P.head_carry + c_end));
    finish_row<OP, false>(P, e.row, t);
}

__device__ __forceinline__ void pusher_send_row(const SpmvParams &P, const PusherParams &Q, uint32_t row) {
    const float v = *reinterpret_cast<volatile const float *>(P.y + row);  // this thread stored it just now
    asm volatile("multimem.st.relaxed.sys.global.f32 [%0], %1;" ::"l"(Q.y_mc + row), "f"(v) : "memory");
}

template <int OP>
__global__ void __launch_bounds__(kPusherThreads, 1) xchg_pusher_kernel(const SpmvParams P, const PusherParams Q) {
    constexpr int U = 4;  // 16-byte loads in flight per thread before the first store
    if (Q.trace && blockIdx.x == 0 && threadIdx.x == 0) Q.trace[0] = glb_now_ns();
    for (uint32_t blk = blockIdx.x; blk < Q.n_blocks; blk += gridDim.x) {
        if (threadIdx.x == 0) {
            const uint32_t first = blk * kPushCtas;
            pusher_spin(Q.push_count + blk, Q.n_ctas - first < kPushCtas ? Q.n_ctas - first : kPushCtas, Q.state + 2);
            Q.push_count[blk] = 0;  // ready for the next step (which starts after this kernel)
            if (Q.trace) Q.trace[4 + 2 * blk] = glb_now_ns();
        }
        __syncthreads();
        for (uint32_t i = Q.blk_fs[blk] + threadIdx.x; i < Q.blk_fs[blk + 1]; i += kPusherThreads) pusher_fix_short<OP>(P, Q.fix_in[i]);
        for (uint32_t i = Q.blk_em[blk] + threadIdx.x; i < Q.blk_em[blk + 1]; i += kPusherThreads)
            finish_row<OP, false>(P, P.empty_rows[i], Semi<OP>::ident());
        __syncthreads();  // the block's rows are final (the loads below go to L2, where those stores are)
        const uint32_t lo = Q.push_lo[blk], hi = Q.push_lo[blk + 1];
        for (uint32_t base = lo & ~3u; base < hi; base += 4u * kPusherThreads * U) {
            float4 v[U];
            uint32_t bits[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const uint32_t r0 = base + 4u * (threadIdx.x + u * kPusherThreads);
                bits[u] = 0;
                if (r0 < hi) {
                    const uint32_t rel = r0 - Q.push_row_base;
                    bits[u] = (__ldg(Q.push_bits + (rel >> 5)) >> (rel & 31u)) & 0xfu;
#pragma unroll
                    for (int i = 0; i < 4; ++i)
                        if (r0 + i < lo || r0 + i >= hi) bits[u] &= ~(1u << i);
                    if (bits[u]) v[u] = __ldcg(reinterpret_cast<const float4 *>(P.y + r0));  // (exchange vectors are padded to 16 bytes)
                }
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const uint32_t r0 = base + 4u * (threadIdx.x + u * kPusherThreads);
                if (bits[u] == 0xfu) {
                    asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(Q.y_mc + r0), "f"(v[u].x),
                                 "f"(v[u].y), "f"(v[u].z), "f"(v[u].w)
                                 : "memory");
                } else if (bits[u]) {
                    const float e[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
#pragma unroll
                    for (int i = 0; i < 4; ++i)
                        if (bits[u] & (1u << i))
                            asm volatile("multimem.st.relaxed.sys.global.f32 [%0], %1;" ::"l"(Q.y_mc + r0 + i), "f"(e[i]) : "memory");
                }
            }
        }
        if (Q.trace) {
            __syncthreads();
            if (threadIdx.x == 0) Q.trace[5 + 2 * blk] = glb_now_ns();
        }
    }
    // every block complete (each pusher CTA has seen all of its blocks): the rows that cross block boundaries
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(Q.pusher_done, 1u);
        pusher_spin(Q.pusher_done, gridDim.x, Q.state + 2);
        if (Q.trace && blockIdx.x == 0) Q.trace[1] = glb_now_ns();
    }
    __syncthreads();
    for (uint32_t i = blockIdx.x * kPusherThreads + threadIdx.x; i < Q.n_fix_def; i += gridDim.x * kPusherThreads) {
        const glb_fixup_t e = Q.fix_def[i];
        pusher_fix_short<OP>(P, e);
        pusher_send_row(P, Q, e.row);
    }
    {
        const unsigned lane = threadIdx.x & 31u;
        const uint32_t n_warps = gridDim.x * (kPusherThreads / 32);
        for (uint32_t i = blockIdx.x * (kPusherThreads / 32) + (threadIdx.x >> 5); i < P.n_fix_long; i += n_warps) {  // warp-uniform
            const glb_fixup_t e = P.fix_long[i];
            const uint32_t c_end = e.c_end & ~GLB_FLAG;
            const bool has_head = (e.c_end & GLB_FLAG) != 0;
            const uint32_t c_stop = has_head ? c_end : c_end + 1;
            float t = Semi<OP>::ident();
            for (uint32_t c = e.c_begin + lane; c < c_stop; c += 32) t = Semi<OP>::add(t, __ldcg(P.tail_carry + c));
#pragma unroll
            for (int d = 16; d >= 1; d >>= 1) t = Semi<OP>::add(t, __shfl_xor_sync(kFull, t, d));
            if (lane == 0) {
                if (has_head) t = Semi<OP>::add(t, __ldcg(P.head_carry + c_end));
                finish_row<OP, false>(P, e.row, t);
                pusher_send_row(P, Q, e.row);
            }
        }
    }
    __threadfence_system();
    __syncthreads();
    if (Q.trace && blockIdx.x == 0 && threadIdx.x == 0) Q.trace[2] = glb_now_ns();
    if (threadIdx.x == 0 && atomicAdd(Q.state + 1, 1u) == gridDim.x - 1) {
        Q.state[1] = 0;
        *Q.pusher_done = 0;
        const uint32_t epoch = Q.state[0] + 1;
        Q.state[0] = epoch;
        // (release: ordered after this CTA's fence above and, through the ticket, after every other CTA's)
        asm volatile("multimem.st.release.sys.global.u32 [%0], %1;" ::"l"(Q.mc_flags + Q.rank), "r"(epoch) : "memory");
        if (Q.trace) Q.trace[3] = glb_now_ns();
    }
}

constexpr int kLaunchMain = 1, kLaunchFixup = 2;

// `st`: the stream of this launch (the context's, or a sub-block stream of a split step); `which`: main kernel,
// fix-up kernel or both.  The main kernel covers chunks [P.chunk_begin, P.chunk_end).
template <int OP, int VT = GLB_VAL_F32>
int launch_op(glb_ctx_t ctx, glb_csr_t m, SpmvParams P, bool *published, cudaStream_t st, int which) {
    const uint32_t n_window = P.chunk_end - P.chunk_begin;
    P.n_ctas = (n_window + kWarpsPerBlock - 1) / kWarpsPerBlock;
    const bool timing = ctx->timing && st == ctx->stream && which == (kLaunchMain | kLaunchFixup);
    cudaEvent_t ev[3] = {nullptr, nullptr, nullptr};
    if (timing) {
        for (auto &e : ev) GLB_CUDA(cudaEventCreate(&e));
        GLB_CUDA(cudaEventRecord(ev[0], st));
    }
    const bool do_main = (which & kLaunchMain) && n_window;
    if (VT == GLB_VAL_F32 && do_main && m->tile_threads) {
        const uint32_t n_warps = m->tile_threads / 32;
        const size_t smem = size_t((P.tile_k + 3u) & ~3u) * 4 + size_t(n_warps) * GLB_ROW_CAP * 4 + 16;
        size_t *attr_set = ctx->tile_smem_set;  // function attributes are per device: cached in the context
        if (attr_set[OP] < smem) {
            GLB_CUDA(cudaFuncSetAttribute(spmv_lane_tile_kernel<OP>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
            attr_set[OP] = smem;
        }
        uint32_t grid = (P.n_chunks + n_warps - 1) / n_warps;  // (no split support: always the whole shard)
        if (grid > uint32_t(ctx->num_sms)) grid = uint32_t(ctx->num_sms);
        spmv_lane_tile_kernel<OP><<<grid, m->tile_threads, smem, st>>>(P);
    } else if (VT == GLB_VAL_F32 && do_main && OP == GLB_OP_LOGICAL_AND_OR && P.xbits) {
        const uint32_t grid = P.n_ctas;
        int &bits_carveout_set = ctx->bits_carveout_set;
        if (bits_carveout_set != m->smem_carveout_pct) {
            const int pct = m->smem_carveout_pct;
            GLB_CUDA(cudaFuncSetAttribute(spmv_lane_bits_kernel<1, false>, cudaFuncAttributePreferredSharedMemoryCarveout, pct));
            GLB_CUDA(cudaFuncSetAttribute(spmv_lane_bits_kernel<1, true>, cudaFuncAttributePreferredSharedMemoryCarveout, pct));
            GLB_CUDA(cudaFuncSetAttribute(spmv_lane_bits_kernel<2, false>, cudaFuncAttributePreferredSharedMemoryCarveout, pct));
            GLB_CUDA(cudaFuncSetAttribute(spmv_lane_bits_kernel<2, true>, cudaFuncAttributePreferredSharedMemoryCarveout, pct));
            bits_carveout_set = pct;
        }
        const bool masked = P.mask_type != GLB_MASK_NONE;
        if (m->all_nonzero && masked) spmv_lane_bits_kernel<2, true><<<grid, kThreads, 0, st>>>(P);
        else if (m->all_nonzero) spmv_lane_bits_kernel<2, false><<<grid, kThreads, 0, st>>>(P);
        else if (masked) spmv_lane_bits_kernel<1, true><<<grid, kThreads, 0, st>>>(P);
        else spmv_lane_bits_kernel<1, false><<<grid, kThreads, 0, st>>>(P);
    } else if (do_main) {
        // leave everything but the staging arrays to L1: that is where the hot x lines live
        int *carveout_set = ctx->carveout_set + 3 * VT;
        if (carveout_set[OP] != m->smem_carveout_pct) {
            GLB_CUDA(cudaFuncSetAttribute(spmv_lane_kernel<OP, false, VT>, cudaFuncAttributePreferredSharedMemoryCarveout,
                                          m->smem_carveout_pct));
            GLB_CUDA(cudaFuncSetAttribute(spmv_lane_kernel<OP, true, VT>, cudaFuncAttributePreferredSharedMemoryCarveout,
                                          m->smem_carveout_pct));
            carveout_set[OP] = m->smem_carveout_pct;
        }
        const uint32_t grid = P.n_ctas;
        if (P.mask_type != GLB_MASK_NONE) spmv_lane_kernel<OP, true, VT><<<grid, kThreads, 0, st>>>(P);
        else spmv_lane_kernel<OP, false, VT><<<grid, kThreads, 0, st>>>(P);
    }
    if (timing) GLB_CUDA(cudaEventRecord(ev[1], st));
    const uint32_t nb_short = (P.n_fix_short + kThreads - 1) / kThreads;
    const uint32_t nb_long = (P.n_fix_long + kWarpsPerBlock - 1) / kWarpsPerBlock;
    const uint32_t nb_empty = (P.n_empty + kThreads - 1) / kThreads;
    if ((which & kLaunchFixup) && nb_short + nb_long + nb_empty) {
        spmv_fixup_kernel<OP, VT><<<nb_short + nb_long + nb_empty, kThreads, 0, st>>>(P, nb_short, nb_long);
        if (published) *published = P.pub_flags_mc != nullptr;
    }
    if (timing) {
        GLB_CUDA(cudaEventRecord(ev[2], st));
        for (auto e : ev) ctx->timing_events.push_back(e);
    }
    GLB_CUDA(cudaGetLastError());
    return GLB_OK;
}

template <typename T>
int upload(glb_ctx_t ctx, T **dptr, const T *host, size_t n, size_t n_alloc, size_t *bytes) {
    *dptr = nullptr;
    if (n_alloc == 0) n_alloc = 1;
    cudaError_t e = cudaMalloc(reinterpret_cast<void **>(dptr), n_alloc * sizeof(T));
    if (e != cudaSuccess) {
        glb_set_error("cudaMalloc(%zu): %s", n_alloc * sizeof(T), cudaGetErrorString(e));
        return GLB_ENOMEM;
    }
    *bytes += n_alloc * sizeof(T);
    if (n_alloc > n) GLB_CUDA(cudaMemsetAsync(*dptr + n, 0, (n_alloc - n) * sizeof(T), ctx->stream));
    if (n) GLB_CUDA(cudaMemcpyAsync(*dptr, host, n * sizeof(T), cudaMemcpyHostToDevice, ctx->stream));
    return GLB_OK;
}

}  // namespace

static int dispatch_op(glb_ctx_t ctx, glb_csr_t m, int op, int val_type, const SpmvParams &P, bool *published, cudaStream_t st,
                       int which) {
    switch (op * 3 + val_type) {
        case GLB_OP_MUL_ADD * 3 + GLB_VAL_F32: return launch_op<GLB_OP_MUL_ADD, GLB_VAL_F32>(ctx, m, P, published, st, which);
        case GLB_OP_LOGICAL_AND_OR * 3 + GLB_VAL_F32: return launch_op<GLB_OP_LOGICAL_AND_OR, GLB_VAL_F32>(ctx, m, P, published, st, which);
        case GLB_OP_ADD_MIN * 3 + GLB_VAL_F32: return launch_op<GLB_OP_ADD_MIN, GLB_VAL_F32>(ctx, m, P, published, st, which);
        case GLB_OP_MUL_ADD * 3 + GLB_VAL_U32: return launch_op<GLB_OP_MUL_ADD, GLB_VAL_U32>(ctx, m, P, published, st, which);
        case GLB_OP_LOGICAL_AND_OR * 3 + GLB_VAL_U32: return launch_op<GLB_OP_LOGICAL_AND_OR, GLB_VAL_U32>(ctx, m, P, published, st, which);
        case GLB_OP_ADD_MIN * 3 + GLB_VAL_U32: return launch_op<GLB_OP_ADD_MIN, GLB_VAL_U32>(ctx, m, P, published, st, which);
        case GLB_OP_MUL_ADD * 3 + GLB_VAL_UFIXED: return launch_op<GLB_OP_MUL_ADD, GLB_VAL_UFIXED>(ctx, m, P, published, st, which);
        case GLB_OP_LOGICAL_AND_OR * 3 + GLB_VAL_UFIXED: return launch_op<GLB_OP_LOGICAL_AND_OR, GLB_VAL_UFIXED>(ctx, m, P, published, st, which);
        case GLB_OP_ADD_MIN * 3 + GLB_VAL_UFIXED: return launch_op<GLB_OP_ADD_MIN, GLB_VAL_UFIXED>(ctx, m, P, published, st, which);
    }
    glb_set_error("glb_spmv: invalid semiring op %d", op);
    return GLB_EINVAL;
}

// Split step of a row-sharded run: the shard's chunks are cut into sub-blocks whose main kernels are launched
// on streams of their own (they run concurrently, like one launch, but each signals its own completion); as
// soon as a sub-block and its fix-up rows are done, its finished rows [sub_row[s], sub_row[s + 1]) go to every
// peer with COPY-ENGINE peer copies, which overlap the sub-blocks still computing without touching the SMs'
// request ports (SM-issued remote stores from inside the kernel measured slower).  The
// caller publishes the epoch after the join.
static int launch_split(glb_ctx_t ctx, glb_csr_t m, int op, int val_type, const SpmvParams &P0, const GlbSpmvSplit &sp) {
    const int S = int(m->sub_chunk.size()) - 1;
    if (!ctx->split_ev_head) {
        GLB_CUDA(cudaEventCreateWithFlags(&ctx->split_ev_head, cudaEventDisableTiming));
        // Earlier sub-blocks get the higher stream priority: the block scheduler then drains them first instead of
        // sharing the SMs evenly between the concurrent launches, so their rows are ready (and travelling) while the
        // later ones still compute.
        int prio_least = 0, prio_greatest = 0;
        GLB_CUDA(cudaDeviceGetStreamPriorityRange(&prio_least, &prio_greatest));
        for (int s = 0; s < GLB_MAX_SPLIT; ++s) {
            int prio = prio_greatest + s;
            if (prio > prio_least) prio = prio_least;
            GLB_CUDA(cudaStreamCreateWithPriority(&ctx->split_stream[s], cudaStreamNonBlocking, prio));
            GLB_CUDA(cudaEventCreateWithFlags(&ctx->split_ev_main[s], cudaEventDisableTiming));
            GLB_CUDA(cudaEventCreateWithFlags(&ctx->split_ev_done[s], cudaEventDisableTiming));
        }
    }
    GLB_CUDA(cudaEventRecord(ctx->split_ev_head, ctx->stream));  // after the head kernel (hot-column pack / bitmap, with the acquire)
    // GLB_XCHG_TRACE=<n>: print the timeline of the n-th split step (device events; debugging, not while recording)
    static const long trace_at = getenv("GLB_XCHG_TRACE") ? atol(getenv("GLB_XCHG_TRACE")) : -1;
    static long call_no = 0;
    const bool trace = trace_at >= 0 && call_no++ == trace_at;
    cudaEvent_t tev[1 + 3 * GLB_MAX_SPLIT] = {};
    if (trace) {
        for (auto &e : tev) cudaEventCreate(&e);
        cudaEventRecord(tev[0], ctx->stream);
    }
    for (int s = 0; s < S; ++s) {
        cudaStream_t st = ctx->split_stream[s];
        GLB_CUDA(cudaStreamWaitEvent(st, ctx->split_ev_head, 0));
        SpmvParams P = P0;
        P.chunk_begin = m->sub_chunk[s];
        P.chunk_end = m->sub_chunk[s + 1];
        int rc = dispatch_op(ctx, m, op, val_type, P, nullptr, st, kLaunchMain);
        if (rc) return rc;
        GLB_CUDA(cudaEventRecord(ctx->split_ev_main[s], st));
        if (trace) cudaEventRecord(tev[1 + 3 * s], st);
        // rows of this sub-block that touch chunk boundaries (their spans may reach back into earlier sub-blocks) and its empty rows
        for (int t = 0; t < s; ++t) GLB_CUDA(cudaStreamWaitEvent(st, ctx->split_ev_main[t], 0));
        P.fix_short = m->fix_short + m->sub_fs[s];
        P.n_fix_short = m->sub_fs[s + 1] - m->sub_fs[s];
        P.fix_long = m->fix_long + m->sub_fl[s];
        P.n_fix_long = m->sub_fl[s + 1] - m->sub_fl[s];
        P.empty_rows = m->empty_rows + m->sub_em[s];
        P.n_empty = m->sub_em[s + 1] - m->sub_em[s];
        rc = dispatch_op(ctx, m, op, val_type, P, nullptr, st, kLaunchFixup);
        if (rc) return rc;
        if (trace) cudaEventRecord(tev[2 + 3 * s], st);
        const uint32_t r0 = m->sub_row[s], r1 = m->sub_row[s + 1];
        // (a copy-engine copy costs 12-17 us whatever its size: fine while other sub-blocks compute, too slow for the
        // LAST one, whose rows the caller sends with the multicast push kernel when there is a multicast mapping)
        for (int p = 0; p < sp.n_peers && r1 > r0 && !(sp.last_by_caller && s == S - 1); ++p)
            GLB_CUDA(cudaMemcpyAsync(sp.y_peers[p] + r0, P0.y + r0, sizeof(float) * size_t(r1 - r0), cudaMemcpyDeviceToDevice, st));
        GLB_CUDA(cudaEventRecord(ctx->split_ev_done[s], st));
        if (trace) cudaEventRecord(tev[3 + 3 * s], st);
    }
    for (int s = 0; s < S; ++s) GLB_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->split_ev_done[s], 0));
    if (trace) {
        cudaDeviceSynchronize();
        for (int s = 0; s < S; ++s) {
            float a = 0, b = 0, c = 0;
            cudaEventElapsedTime(&a, tev[0], tev[1 + 3 * s]);
            cudaEventElapsedTime(&b, tev[0], tev[2 + 3 * s]);
            cudaEventElapsedTime(&c, tev[0], tev[3 + 3 * s]);
            fprintf(stderr, "[glb split] sub-block %d: chunks [%u, %u) rows [%u, %u) main done +%.1f us, fix-up done +%.1f us, copies done +%.1f us\n",
                    s, m->sub_chunk[s], m->sub_chunk[s + 1], m->sub_row[s], m->sub_row[s + 1], a * 1e3, b * 1e3, c * 1e3);
        }
        for (auto &e : tev) cudaEventDestroy(e);
    }
    return GLB_OK;
}

bool glb_pusher_applies(glb_csr_t m, const float *y, const float *y_mc) {
    const bool aligned16 = ((reinterpret_cast<uintptr_t>(y) | reinterpret_cast<uintptr_t>(y_mc)) & 15u) == 0;
    return m->pusher_bits && m->n_chunks && !m->tile_threads && aligned16;
}

// Fork: the pusher kernel starts with the step (its CTAs are resident before the main kernel fills the machine);
// the caller joins ctx->pusher_ev_done into ctx->stream after the step's SpMV kernels.
//   GLB_XCHG_PUSHERS=<n>         pusher CTAs (default 4)
//   GLB_XCHG_PUSHER_SMEM_KB=<n>  shared memory each one requests (default 227 = the whole SM; 0: share the SM with compute CTAs)
template <int OP>
static int launch_pusher_op(glb_ctx_t ctx, const SpmvParams &P, const PusherParams &Q, uint32_t grid, int smem) {
    if (ctx->pusher_smem_set[OP] != smem) {
        GLB_CUDA(cudaFuncSetAttribute(xchg_pusher_kernel<OP>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        if (smem) GLB_CUDA(cudaFuncSetAttribute(xchg_pusher_kernel<OP>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
        ctx->pusher_smem_set[OP] = smem;
    }
    xchg_pusher_kernel<OP><<<grid, kPusherThreads, smem, ctx->pusher_stream>>>(P, Q);
    GLB_CUDA(cudaGetLastError());
    return GLB_OK;
}

int glb_launch_pusher(glb_ctx_t ctx, glb_csr_t m, int op, float zero, int mask_type, const float *mask, float *y,
                      const glb_spmv_epilogue_t *ep, float *y_mc, uint32_t *mc_flags, uint32_t *state, int rank) {
    auto env_u = [](const char *name, unsigned dflt) { const char *v = getenv(name); return v && *v ? unsigned(strtoul(v, nullptr, 10)) : dflt; };
    static const uint32_t n_pushers = std::max(1u, std::min(32u, env_u("GLB_XCHG_PUSHERS", 4)));
    static const int smem = int(std::min(227u, env_u("GLB_XCHG_PUSHER_SMEM_KB", 227))) * 1024;
    if (!ctx->pusher_stream) {
        // CUDA loads kernels lazily, and loading one can wait for running kernels to finish: a pusher kernel spinning on
        // counters of an SpMV kernel that is still to be LOADED would never see them move.  Load everything a step can
        // launch before the first pusher goes out.
        cudaFuncAttributes fa;
#define GLB_PRELOAD(k) GLB_CUDA(cudaFuncGetAttributes(&fa, k))
#define GLB_PRELOAD_OP(OP)                                           \
        GLB_PRELOAD((spmv_lane_kernel<OP, false, GLB_VAL_F32>));     \
        GLB_PRELOAD((spmv_lane_kernel<OP, true, GLB_VAL_F32>));      \
        GLB_PRELOAD((spmv_fixup_kernel<OP, GLB_VAL_F32>));           \
        GLB_PRELOAD(xchg_pusher_kernel<OP>)
        GLB_PRELOAD_OP(GLB_OP_MUL_ADD);
        GLB_PRELOAD_OP(GLB_OP_LOGICAL_AND_OR);
        GLB_PRELOAD_OP(GLB_OP_ADD_MIN);
        GLB_PRELOAD((spmv_lane_bits_kernel<1, false>));
        GLB_PRELOAD((spmv_lane_bits_kernel<1, true>));
        GLB_PRELOAD((spmv_lane_bits_kernel<2, false>));
        GLB_PRELOAD((spmv_lane_bits_kernel<2, true>));
        GLB_PRELOAD(gather_hot_kernel);
        GLB_PRELOAD(pack_bits_kernel);
#undef GLB_PRELOAD_OP
#undef GLB_PRELOAD
        int rc = glb_xchg_preload();
        if (rc) return rc;
        int prio_least = 0, prio_greatest = 0;
        GLB_CUDA(cudaDeviceGetStreamPriorityRange(&prio_least, &prio_greatest));
        GLB_CUDA(cudaStreamCreateWithPriority(&ctx->pusher_stream, cudaStreamNonBlocking, prio_greatest));
        GLB_CUDA(cudaEventCreateWithFlags(&ctx->pusher_ev_fork, cudaEventDisableTiming));
        GLB_CUDA(cudaEventCreateWithFlags(&ctx->pusher_ev_done, cudaEventDisableTiming));
    }
    SpmvParams P;  // what finish_row and the fix-ups read
    memset(&P, 0, sizeof(P));
    P.y = y;
    P.mask = mask;
    P.zero = zero;
    P.mask_type = mask_type;
    if (ep) {
        P.add_enable = ep->add_enable;
        P.add_val = ep->add_val;
        P.assign_inout = ep->assign_inout;
        P.assign_val = ep->assign_val;
        P.assign_mask_type = ep->assign_mask_type;
    }
    P.head_carry = m->head_carry;
    P.tail_carry = m->tail_carry;
    P.empty_rows = m->empty_rows;
    P.fix_long = m->fix_long;
    P.n_fix_long = m->n_fix_long;
    PusherParams Q;
    memset(&Q, 0, sizeof(Q));
    Q.y_mc = y_mc;
    Q.push_bits = m->pusher_bits;
    Q.push_lo = m->push_lo;
    Q.push_count = m->push_count;
    Q.pusher_done = m->push_count + m->n_push_blocks;
    Q.fix_in = m->fix_in;
    Q.fix_def = m->fix_def;
    Q.blk_fs = m->blk_fs;
    Q.blk_em = m->blk_em;
    Q.n_fix_def = m->n_fix_def;
    Q.push_row_base = m->row_begin & ~31u;
    Q.n_blocks = m->n_push_blocks;
    Q.n_ctas = (m->n_chunks + kWarpsPerBlock - 1) / kWarpsPerBlock;
    Q.mc_flags = mc_flags;
    Q.state = state;
    Q.rank = rank;
    // GLB_XCHG_TRACE=<n>: timeline of the n-th pusher launch of the process (eager launches only), printed at the next one
    static const long trace_at = getenv("GLB_XCHG_TRACE") ? atol(getenv("GLB_XCHG_TRACE")) : -1;
    static long call_no = 0;
    static unsigned long long *d_trace = nullptr;
    if (trace_at >= 0) {
        const size_t n_stamps = 4 + 2 * size_t(m->n_push_blocks);
        if (call_no == trace_at) {
            GLB_CUDA(cudaMalloc(reinterpret_cast<void **>(&d_trace), n_stamps * sizeof(unsigned long long)));
            GLB_CUDA(cudaMemset(d_trace, 0, n_stamps * sizeof(unsigned long long)));
            Q.trace = d_trace;
        } else if (call_no == trace_at + 1 && d_trace) {
            std::vector<unsigned long long> t(n_stamps);
            GLB_CUDA(cudaDeviceSynchronize());
            GLB_CUDA(cudaMemcpy(t.data(), d_trace, n_stamps * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
            auto us = [&](size_t i) { return t[i] ? double(static_cast<long long>(t[i] - t[0])) * 1e-3 : -1.0; };
            fprintf(stderr, "[glb pusher rank %d] %u blocks; all blocks out +%.1f us, late rows fenced +%.1f us, published +%.1f us\n", rank,
                    m->n_push_blocks, us(1), us(2), us(3));
            for (uint32_t b = 0; b < m->n_push_blocks; b += std::max(1u, m->n_push_blocks / 16))
                fprintf(stderr, "[glb pusher rank %d]   block %u ready +%.1f us, sent +%.1f us\n", rank, b, us(4 + 2 * b), us(5 + 2 * b));
            const uint32_t lastb = m->n_push_blocks - 1;
            fprintf(stderr, "[glb pusher rank %d]   block %u ready +%.1f us, sent +%.1f us\n", rank, lastb, us(4 + 2 * lastb), us(5 + 2 * lastb));
        }
        ++call_no;
    }
    GLB_CUDA(cudaEventRecord(ctx->pusher_ev_fork, ctx->stream));
    GLB_CUDA(cudaStreamWaitEvent(ctx->pusher_stream, ctx->pusher_ev_fork, 0));
    const uint32_t grid = std::min(n_pushers, std::max(1u, m->n_push_blocks));
    int rc = GLB_EINVAL;
    switch (op) {
        case GLB_OP_MUL_ADD: rc = launch_pusher_op<GLB_OP_MUL_ADD>(ctx, P, Q, grid, smem); break;
        case GLB_OP_LOGICAL_AND_OR: rc = launch_pusher_op<GLB_OP_LOGICAL_AND_OR>(ctx, P, Q, grid, smem); break;
        case GLB_OP_ADD_MIN: rc = launch_pusher_op<GLB_OP_ADD_MIN>(ctx, P, Q, grid, smem); break;
    }
    if (rc) return rc;
    GLB_CUDA(cudaEventRecord(ctx->pusher_ev_done, ctx->pusher_stream));
    return GLB_OK;
}

int glb_launch_spmv(glb_ctx_t ctx, glb_csr_t m, int op, float zero, int mask_type, const float *x, const float *mask,
                    float *y, const glb_spmv_epilogue_t *ep, float *const *y_peers, int n_peers, const GlbSpmvMc *mc,
                    const GlbXchgWait *wait, bool *published, int val_type, const GlbSpmvSplit *split) {
    SpmvParams P;
    memset(&P, 0, sizeof(P));
    if (published) *published = false;
    float *y_mc = mc ? mc->y_mc : nullptr;
    const bool aligned16 = ((reinterpret_cast<uintptr_t>(y) | reinterpret_cast<uintptr_t>(y_mc)) & 15u) == 0;
    if (mc && mc->progressive && m->push_bits && m->n_chunks && !m->tile_threads && aligned16) {
        P.push_bits = m->push_bits;
        P.push_lo = m->push_lo;
        P.push_count = m->push_count;
        P.push_row_base = m->row_begin & ~31u;
        P.pub_flags_mc = mc->pub_flags_mc;
        P.pub_state = mc->pub_state;
        P.pub_rank = mc->rank;
    } else if (mc && mc->pusher) {  // (the caller checked glb_pusher_applies and launched the pusher kernel)
        P.push_bits = m->push_bits;
        P.push_lo = m->push_lo;
        P.push_count = m->push_count;
        P.push_row_base = m->row_begin & ~31u;
        P.push_count_only = 1;
        y_mc = nullptr;  // no kernel of this launch stores to the multicast address
    } else if (mc) {
        P.mc_rows_in_main = 1;
    }
    GlbXchgWait w;
    memset(&w, 0, sizeof(w));
    if (wait) w = *wait;
    const bool head_is_pack = val_type == GLB_VAL_F32 && op == GLB_OP_LOGICAL_AND_OR && m->xbits && m->n_chunks && !m->tile_threads;
    const bool head_is_gather = !head_is_pack && m->n_hot && m->n_chunks;
    if (wait && !head_is_pack && !head_is_gather) {  // no kernel of this launch opens with the acquire: its own launch
        int rc = glb_xchg_wait_launch(ctx, w);
        if (rc) return rc;
    }
    P.y_mc = y_mc;
    P.n_peers = n_peers;
    for (int p = 0; p < n_peers && p < GLB_MAX_PEERS; ++p) P.y_peer[p] = y_peers[p];
    P.stream = m->stream;
    P.flags = m->flags;
    P.chunk_goff = m->chunk_goff;
    P.chunk_first = m->chunk_first;
    P.nz_rows = m->nz_rows;
    P.tile_k = m->tile_k;
    if (head_is_pack) {
        // or-and: one bit per stored column word replaces the fp32 gathers
        const uint32_t n_words32 = (m->tile_k + m->num_cols + 31u) / 32u;
        //   GLB_SPMV_MASK_BITS=0   masked or-and launches read the fp32 mask (default 1: a bitmap packed with x)
        static const bool mask_bits = !getenv("GLB_SPMV_MASK_BITS") || atoi(getenv("GLB_SPMV_MASK_BITS")) != 0;
        const uint32_t mbase = m->row_begin & ~31u;
        const uint32_t n_mwords = (mask_type != GLB_MASK_NONE && mask_bits && m->mbits) ? (m->row_end - mbase + 31u) / 32u : 0u;
        const uint32_t threads = (n_words32 + n_mwords) * 32u;
        pack_bits_kernel<<<(threads + kThreads - 1) / kThreads, kThreads, 0, ctx->stream>>>(
            x, m->hot_cols, m->xbits, m->tile_k, m->n_hot, m->num_cols, n_words32, w, mask, m->mbits, mbase, m->row_begin, m->row_end, n_mwords);
        if (n_mwords) {
            P.mbits = m->mbits;
            P.mbits_base = mbase;
        }
        P.xbits = m->xbits;
        P.hot_x = x;
    } else if (m->n_hot && m->n_chunks) {  // pack the x values of the hot columns
        gather_hot_kernel<<<(m->n_hot + kThreads - 1) / kThreads, kThreads, 0, ctx->stream>>>(x, m->hot_cols, m->hot_x,
                                                                                            m->n_hot, w);
        P.hot_x = m->hot_x;
    } else {
        P.hot_x = x;  // tile_k == num_cols (every column hot, identity numbering) or tile_k == 0
    }
    P.x_cold = x - m->tile_k;
    P.mask = mask;
    P.y = y;
    P.head_carry = m->head_carry;
    P.tail_carry = m->tail_carry;
    P.n_chunks = m->n_chunks;
    P.chunk_begin = 0;
    P.chunk_end = m->n_chunks;
    P.zero = zero;
    P.mask_type = mask_type;
    if (ep) {
        P.add_enable = ep->add_enable;
        P.add_val = ep->add_val;
        P.assign_inout = ep->assign_inout;
        P.assign_val = ep->assign_val;
        P.assign_mask_type = ep->assign_mask_type;
    }
    P.fix_short = m->fix_short;
    P.fix_long = m->fix_long;
    P.empty_rows = m->empty_rows;
    P.n_fix_short = m->n_fix_short;
    P.n_fix_long = m->n_fix_long;
    P.n_empty = m->n_empty;
    P.n_nz_rows = m->n_nz_rows;
    static const bool shortcuts = !getenv("GLB_SPMV_SHORTCUTS") || atoi(getenv("GLB_SPMV_SHORTCUTS")) != 0;
    if (shortcuts) {
        P.rows_identity = m->n_empty == 0 && m->n_nz_rows == m->row_end - m->row_begin;
        P.row0 = m->row_begin;
        P.uniform_groups = m->uniform_groups;
        P.last_groups = m->last_groups;
    }
    if (P.push_count_only) return dispatch_op(ctx, m, op, val_type, P, published, ctx->stream, kLaunchMain);  // the pusher CTAs do the fix-ups
    if (!split) return dispatch_op(ctx, m, op, val_type, P, published, ctx->stream, kLaunchMain | kLaunchFixup);
    return launch_split(ctx, m, op, val_type, P, *split);
}

// ------------------------------------------------------------------------------------------
// Host side of the formatter (the counterpart of csr2cpsr, data_formatter.h:457-534).  Shared by
// glb_csr_create and the host-only glb_csr_format_host, which lets the layout be checked
// without a GPU.  O(nnz) work; the stream transposition is spread over the host cores.
struct HostLayout {
    uint32_t *stream = nullptr;  // malloc'd, 256 * n_groups words
    std::vector<uint32_t> flags, chunk_goff, chunk_first, nz_rows, empty_rows, hot_cols;
    std::vector<glb_fixup_t> fix_short, fix_long;
    uint64_t nnz = 0, sb = 0;
    uint32_t n_chunks = 0, n_groups = 0, tile_k = 0, max_groups = GLB_MAX_GROUPS;
    bool all_nonzero = false;  // no stored value is 0.0f: or-and may skip the value stream
    ~HostLayout() { free(stream); }
};

template <typename F>
static void parallel_for(size_t n, size_t grain, F &&body) {
    unsigned hw = std::thread::hardware_concurrency();
    size_t workers = std::min<size_t>(hw ? hw : 1, (n + grain - 1) / grain);
    if (workers <= 1) { body(size_t(0), n); return; }
    std::vector<std::thread> pool;
    const size_t per = (n + workers - 1) / workers;
    for (size_t w = 0; w < workers; ++w) {
        const size_t b = w * per, e = std::min(n, b + per);
        if (b >= e) break;
        pool.emplace_back([&body, b, e] { body(b, e); });
    }
    for (auto &t : pool) t.join();
}

static int format_host(uint32_t num_rows, uint32_t num_cols, const uint32_t *indptr, const uint32_t *indices,
                       const float *data, uint32_t row_begin, uint32_t row_end, uint32_t tile_k_req, HostLayout &L) {
    GLB_REQUIRE(indptr, "NULL indptr");
    GLB_REQUIRE(row_begin <= row_end && row_end <= num_rows, "bad row range");
    GLB_REQUIRE(num_cols >= 1 && num_cols < GLB_FLAG, "num_cols must be in [1, 2^31)");
    const uint64_t sb = indptr[row_begin];
    GLB_REQUIRE(indptr[row_end] >= sb, "indptr not monotone");
    const uint64_t nnz = uint64_t(indptr[row_end]) - sb;
    GLB_REQUIRE(nnz == 0 || indices, "NULL indices");
    L.nnz = nnz;
    L.sb = sb;
    const uint32_t *ix = indices + sb;

    // rows: non-empty (ordinal order) and empty ones
    L.nz_rows.reserve(row_end - row_begin);
    for (uint32_t r = row_begin; r < row_end; ++r) {
        if (indptr[r + 1] < indptr[r]) {
            glb_set_error("glb_csr_create: indptr not monotone at row %u", r);
            return GLB_EINVAL;
        }
        if (indptr[r + 1] == indptr[r]) L.empty_rows.push_back(r);
        else L.nz_rows.push_back(r);
    }
    const uint32_t n_nz = uint32_t(L.nz_rows.size());

    // ---- hot columns: the tile_k most referenced ones get the low numbers ------------------
    std::vector<uint32_t> count(num_cols, 0);
    for (uint64_t i = 0; i < nnz; ++i) {
        if (ix[i] >= num_cols) {
            glb_set_error("glb_csr_create: column index %u out of range at nnz %llu", ix[i], (unsigned long long)(sb + i));
            return GLB_EINVAL;
        }
        count[ix[i]]++;
    }
    std::vector<uint32_t> enc;  // column -> stored word (empty = identity)
    if (tile_k_req >= num_cols) {
        L.tile_k = num_cols;  // every column hot, identity numbering: hot_x is x itself
    } else if (tile_k_req == 0 || nnz == 0) {
        L.tile_k = 0;
    } else {
        L.tile_k = tile_k_req;
        std::vector<uint32_t> order(num_cols);
        for (uint32_t c = 0; c < num_cols; ++c) order[c] = c;
        auto hotter = [&](uint32_t a, uint32_t b) { return count[a] != count[b] ? count[a] > count[b] : a < b; };
        std::nth_element(order.begin(), order.begin() + L.tile_k, order.end(), hotter);
        std::sort(order.begin(), order.begin() + L.tile_k, hotter);
        L.hot_cols.assign(order.begin(), order.begin() + L.tile_k);
        enc.resize(num_cols);
        for (uint32_t c = 0; c < num_cols; ++c) enc[c] = L.tile_k + c;
        for (uint32_t k = 0; k < L.tile_k; ++k) enc[L.hot_cols[k]] = k;
    }
    count.clear();
    count.shrink_to_fit();

    // ---- chunking: up to 8 groups of 128 non-zeros, at most GLB_ROW_CAP row ends ---------------
    // Chunk size: 8 groups (1024 non-zeros) amortise the per-chunk scan and write-back best, but one
    // warp owns one chunk and a B200 holds 148 SMs x 5 CTAs x 8 warps = 5920 of them at a time: a shard
    // of a few million non-zeros (the BFS graph; 1/8 of C2 on 8 GPUs) would run 2-3 waves and lose up
    // to a third of the machine to the last, partly filled wave.  Small shards therefore get smaller
    // chunks.  Measured on the BFS graph (13 M non-zeros): 256-nnz chunks (11 waves) ran a pull level in
    // 45 us against 25 us with 1024-nnz chunks (2.2 waves) -- the per-chunk scan, flag words and
    // fix-up rows cost more than the partly filled last wave -- so the threshold is low: only shards
    // that would not even fill half the machine are cut finer.  GLB_SPMV_MAX_GROUPS=<1..8> overrides.
    uint32_t max_groups = GLB_MAX_GROUPS;
    {
        const uint64_t want_chunks = GLB_SPMV_WANT_CHUNKS;
        while (max_groups > 1 && nnz / (uint64_t(GLB_GROUP) * max_groups) < want_chunks) max_groups >>= 1;
        if (const char *v = getenv("GLB_SPMV_MAX_GROUPS")) {
            const unsigned long g = strtoul(v, nullptr, 10);
            if (g >= 1 && g <= GLB_MAX_GROUPS) max_groups = uint32_t(g);
        }
    }
    L.max_groups = max_groups;
    const uint64_t kChunkMax = uint64_t(GLB_GROUP) * max_groups;
    std::vector<uint64_t> cs;  // chunk start positions (shard-local), plus the end sentinel
    {
        uint64_t cur = 0;
        uint32_t flags_in = 0;
        for (uint32_t k = 0; k < n_nz; ++k) {
            const uint64_t p0 = uint64_t(indptr[L.nz_rows[k]]) - sb;
            while (p0 >= cur + kChunkMax) { cs.push_back(cur); cur += kChunkMax; flags_in = 0; }
            if (p0 == cur) continue;  // the row opens the chunk: no flag
            if (flags_in == GLB_ROW_CAP - 1) {  // keep one slot for the padding flag: cut here, at a row start
                cs.push_back(cur);
                cur = p0;
                flags_in = 0;
                continue;
            }
            ++flags_in;
        }
        while (cur < nnz) { cs.push_back(cur); cur = std::min(nnz, cur + kChunkMax); }
        cs.push_back(nnz);
    }
    const uint32_t n_chunks = uint32_t(cs.size() - 1);
    L.n_chunks = n_chunks;
    L.chunk_goff.assign(n_chunks + 1, 0);
    for (uint32_t c = 0; c < n_chunks; ++c)
        L.chunk_goff[c + 1] = L.chunk_goff[c] + uint32_t((cs[c + 1] - cs[c] + GLB_GROUP - 1) / GLB_GROUP);
    L.n_groups = L.chunk_goff[n_chunks];
    auto padded = [&](uint32_t c) { return ((cs[c + 1] - cs[c]) % GLB_GROUP) != 0; };
    auto n_of = [&](uint32_t c) { return L.chunk_goff[c + 1] - L.chunk_goff[c]; };

    // ---- flags, chunk_first, fix-up list -----------------------------------------------------
    L.flags.assign(size_t(n_chunks) * 32, 0);
    L.chunk_first.assign(n_chunks, 0);
    auto set_flag = [&](uint32_t c, uint64_t i) {  // position i of chunk c -> lane word, bit
        const uint32_t run = 4 * n_of(c);
        L.flags[size_t(c) * 32 + i / run] |= 1u << (i % run);
    };
    {
        uint32_t c = 0;
        for (uint32_t k = 0; k < n_nz; ++k) {
            const uint32_t r = L.nz_rows[k];
            const uint64_t p0 = uint64_t(indptr[r]) - sb, p1 = uint64_t(indptr[r + 1]) - sb;
            while (cs[c + 1] <= p0) ++c;
            const uint32_t c_s = c;
            if (p0 == cs[c_s]) L.chunk_first[c_s] = k | GLB_FLAG;  // row starts with the chunk
            else set_flag(c_s, p0 - cs[c_s]);
            uint32_t c_e = c_s;
            while (cs[c_e + 1] < p1) { ++c_e; L.chunk_first[c_e] = k; }  // chunks starting inside this row
            const bool closed = (p1 < cs[c_e + 1]) || padded(c_e);  // a later flag of chunk c_e ends the row
            if (c_s == c_e && closed) continue;  // interior row: the main kernel finishes it
            glb_fixup_t e;
            e.row = r;
            e.c_begin = c_s;
            e.c_end = closed ? (c_e | GLB_FLAG) : c_e;
            ((c_e - c_s) <= 32 ? L.fix_short : L.fix_long).push_back(e);
        }
        for (uint32_t c2 = 0; c2 < n_chunks; ++c2)
            if (padded(c2)) set_flag(c2, cs[c2 + 1] - cs[c2]);  // padding is a row of its own, never written back
    }

    // ---- the transposed stream -----------------------------------------------------------------
    L.stream = static_cast<uint32_t *>(malloc(sizeof(uint32_t) * 256 * size_t(L.n_groups ? L.n_groups : 1)));
    if (!L.stream) { glb_set_error("glb_csr_create: host allocation failed"); return GLB_ENOMEM; }
    const uint32_t *dbits = reinterpret_cast<const uint32_t *>(data ? data + sb : nullptr);
    L.all_nonzero = false;
    if (data) {
        bool any_zero = false;
        for (uint64_t i = 0; i < nnz && !any_zero; ++i) any_zero = (data[sb + i] == 0.0f);
        L.all_nonzero = !any_zero;
    }
    const uint32_t *encp = enc.empty() ? nullptr : enc.data();
    parallel_for(n_chunks, 256, [&](size_t cb, size_t ce) {
        for (size_t c = cb; c < ce; ++c) {
            const uint32_t n = L.chunk_goff[c + 1] - L.chunk_goff[c], run = 4 * n;
            uint32_t *out = L.stream + size_t(L.chunk_goff[c]) * 256;
            const uint64_t s = cs[c], len = cs[c + 1] - cs[c];
            uint64_t i = 0;
            for (uint32_t lane = 0; lane < 32; ++lane) {
                for (uint32_t r = 0; r < run; ++r, ++i) {
                    const uint32_t w = (r >> 2) * 256 + lane * 4 + (r & 3);
                    if (i < len) {
                        const uint32_t col = ix[s + i];
                        out[w] = encp ? encp[col] : col;
                        out[w + 128] = dbits ? dbits[s + i] : 0u;
                    } else {
                        out[w] = 0;  // padding: a valid (hot) address, value irrelevant
                        out[w + 128] = 0;
                    }
                }
            }
        }
    });
    return GLB_OK;
}

extern "C" {

// Host-only view of the layout for tests (no CUDA call).  Arrays are malloc'd copies the
// caller releases with glb_host_layout_free.
int glb_csr_format_host(uint32_t num_rows, uint32_t num_cols, const uint32_t *indptr, const uint32_t *indices,
                        const float *data, uint32_t row_begin, uint32_t row_end, uint32_t tile_k,
                        glb_host_layout_t *out) {
    GLB_REQUIRE(out, "out is NULL");
    memset(out, 0, sizeof(*out));
    HostLayout L;
    int rc = format_host(num_rows, num_cols, indptr, indices, data, row_begin, row_end, tile_k, L);
    if (rc) return rc;
    auto dup = [](const void *src, size_t bytes) {
        void *p = malloc(bytes ? bytes : 1);
        if (p && bytes) memcpy(p, src, bytes);
        return static_cast<uint32_t *>(p);
    };
    std::vector<glb_fixup_t> fix(L.fix_short);
    fix.insert(fix.end(), L.fix_long.begin(), L.fix_long.end());
    out->group = GLB_GROUP;
    out->max_groups = L.max_groups;
    out->row_cap = GLB_ROW_CAP;
    out->nnz = L.nnz;
    out->n_chunks = L.n_chunks;
    out->n_groups = L.n_groups;
    out->n_nz_rows = uint32_t(L.nz_rows.size());
    out->n_empty = uint32_t(L.empty_rows.size());
    out->n_fixups = uint32_t(fix.size());
    out->tile_k = L.tile_k;
    out->n_hot = uint32_t(L.hot_cols.size());
    out->stream = dup(L.stream, sizeof(uint32_t) * 256 * size_t(L.n_groups));
    out->flags = dup(L.flags.data(), sizeof(uint32_t) * L.flags.size());
    out->chunk_goff = dup(L.chunk_goff.data(), sizeof(uint32_t) * L.chunk_goff.size());
    out->chunk_first = dup(L.chunk_first.data(), sizeof(uint32_t) * L.chunk_first.size());
    out->nz_rows = dup(L.nz_rows.data(), sizeof(uint32_t) * L.nz_rows.size());
    out->empty_rows = dup(L.empty_rows.data(), sizeof(uint32_t) * L.empty_rows.size());
    out->fixups = dup(fix.data(), sizeof(glb_fixup_t) * fix.size());
    out->hot_cols = dup(L.hot_cols.data(), sizeof(uint32_t) * L.hot_cols.size());
    return GLB_OK;
}

int glb_host_layout_free(glb_host_layout_t *l) {
    if (!l) return GLB_OK;
    free(l->stream); free(l->flags); free(l->chunk_goff); free(l->chunk_first);
    free(l->nz_rows); free(l->empty_rows); free(l->fixups); free(l->hot_cols);
    memset(l, 0, sizeof(*l));
    return GLB_OK;
}

int glb_csr_create(glb_ctx_t ctx, uint32_t num_rows, uint32_t num_cols, const uint32_t *indptr, const uint32_t *indices,
                   const float *data, uint32_t row_begin, uint32_t row_end, glb_csr_t *out) {
    GLB_REQUIRE(ctx && out, "NULL argument");
    *out = nullptr;
    // Tuning knobs (defaults chosen by measurement on B200):
    //   GLB_SPMV_TILE_K=<floats>     hot columns packed into the L1-resident hot_x (0 = none)
    //   GLB_SPMV_CARVEOUT=<percent>  shared-memory carve-out hint of the main kernel
    auto env_u32 = [](const char *name, uint32_t dflt) {
        const char *v = getenv(name);
        return v ? uint32_t(strtoul(v, nullptr, 10)) : dflt;
    };
    const uint32_t tile_k = env_u32("GLB_SPMV_TILE_K", GLB_DEFAULT_TILE_K);
    const uint32_t carveout = env_u32("GLB_SPMV_CARVEOUT", GLB_DEFAULT_CARVEOUT_PCT);
    //   GLB_SPMV_TILE_THREADS=<n>    0: hot vector through L1 (one chunk per warp); 32..1024: persistent
    //                                CTAs of n threads that keep it in shared memory (TMA bulk load)
    uint32_t tile_threads = env_u32("GLB_SPMV_TILE_THREADS", GLB_DEFAULT_TILE_THREADS);
    if (tile_threads > 1024 || tile_threads % 32) tile_threads = 1024;

    HostLayout L;
    int rc = format_host(num_rows, num_cols, indptr, indices, data, row_begin, row_end, tile_k, L);
    if (rc) return rc;
    GLB_REQUIRE(L.nnz == 0 || data, "NULL data");
    GLB_CUDA(cudaSetDevice(ctx->device));

    glb_csr_t m = new glb_csr_s();
    m->ctx = ctx;
    glb_ctx_retain(ctx);
    m->num_rows = num_rows;
    m->num_cols = num_cols;
    m->row_begin = row_begin;
    m->row_end = row_end;
    m->nnz = L.nnz;
    m->n_chunks = L.n_chunks;
    m->n_groups = L.n_groups;
    m->n_nz_rows = uint32_t(L.nz_rows.size());
    m->n_fix_short = uint32_t(L.fix_short.size());
    m->n_fix_long = uint32_t(L.fix_long.size());
    m->n_empty = uint32_t(L.empty_rows.size());
    m->tile_k = L.tile_k;
    m->n_hot = uint32_t(L.hot_cols.size());
    m->smem_carveout_pct = int(carveout > 100 ? 100 : carveout);
    m->all_nonzero = L.all_nonzero;
    // the shared-memory flavour needs the tile, the staging arrays and the barrier in 227 KB
    if (L.tile_k == 0 || size_t((L.tile_k + 3u) & ~3u) * 4 + size_t(tile_threads / 32) * GLB_ROW_CAP * 4 + 16 > 227 * 1024)
        tile_threads = 0;
    m->tile_threads = tile_threads;

    size_t bytes = 0;
    const size_t words = size_t(L.n_groups) * 256;
    if (!rc) rc = upload(ctx, &m->stream, L.stream, words, words, &bytes);
    if (!rc) rc = upload(ctx, &m->flags, L.flags.data(), L.flags.size(), L.flags.size(), &bytes);
    if (!rc) rc = upload(ctx, &m->chunk_goff, L.chunk_goff.data(), L.chunk_goff.size(), L.chunk_goff.size(), &bytes);
    if (!rc) rc = upload(ctx, &m->chunk_first, L.chunk_first.data(), L.chunk_first.size(), L.chunk_first.size(), &bytes);
    if (!rc) rc = upload(ctx, &m->nz_rows, L.nz_rows.data(), L.nz_rows.size(), L.nz_rows.size(), &bytes);
    if (!rc) rc = upload(ctx, &m->fix_short, L.fix_short.data(), L.fix_short.size(), L.fix_short.size(), &bytes);
    if (!rc) rc = upload(ctx, &m->fix_long, L.fix_long.data(), L.fix_long.size(), L.fix_long.size(), &bytes);
    if (!rc) rc = upload(ctx, &m->empty_rows, L.empty_rows.data(), L.empty_rows.size(), L.empty_rows.size(), &bytes);
    if (!rc && m->n_hot) {
        rc = upload(ctx, &m->hot_cols, L.hot_cols.data(), L.hot_cols.size(), L.hot_cols.size(), &bytes);
        if (!rc) rc = upload<float>(ctx, &m->hot_x, nullptr, 0, m->n_hot, &bytes);
    }
    //   GLB_SPMV_BITS=0              or-and SpMV gathers fp32 x like the other semirings (default 1: bitmap)
    if (!rc && env_u32("GLB_SPMV_BITS", 1) && L.n_chunks)
        rc = upload<uint32_t>(ctx, &m->xbits, nullptr, 0, (size_t(L.tile_k) + num_cols + 31) / 32 + 1, &bytes);
    if (!rc && m->xbits) rc = upload<uint32_t>(ctx, &m->mbits, nullptr, 0, (size_t(row_end - (row_begin & ~31u)) + 31) / 32 + 1, &bytes);
    {
        bool uniform = L.n_chunks > 0;
        for (uint32_t c = 0; uniform && c + 1 < L.n_chunks; ++c) uniform = L.chunk_goff[c + 1] - L.chunk_goff[c] == L.max_groups;
        m->uniform_groups = uniform ? L.max_groups : 0;
        m->last_groups = uniform ? L.chunk_goff[L.n_chunks] - L.chunk_goff[L.n_chunks - 1] : 0;
    }
    if (!rc && L.n_chunks) {
        // progressive push of a row-sharded run (push_block_when_complete): which rows the main kernel
        // writes, and the row boundaries of the blocks of kPushCtas consecutive CTAs
        const uint32_t base = row_begin & ~31u;
        std::vector<uint32_t> bits((size_t(row_end - base) + 31) / 32 + 1, 0u);
        for (uint32_t r : L.nz_rows) bits[(r - base) >> 5] |= 1u << ((r - base) & 31u);
        for (const auto &e : L.fix_short) bits[(e.row - base) >> 5] &= ~(1u << ((e.row - base) & 31u));
        for (const auto &e : L.fix_long) bits[(e.row - base) >> 5] &= ~(1u << ((e.row - base) & 31u));
        const uint32_t n_ctas = (L.n_chunks + kWarpsPerBlock - 1) / kWarpsPerBlock;
        const uint32_t n_blocks = (n_ctas + kPushCtas - 1) / kPushCtas;
        std::vector<uint32_t> lo(n_blocks + 1, row_end);
        for (uint32_t j = 0; j < n_blocks; ++j) {
            const uint32_t k = L.chunk_first[size_t(j) * kPushCtas * kWarpsPerBlock] & ~GLB_FLAG;
            lo[j] = k < L.nz_rows.size() ? L.nz_rows[k] : row_end;
        }
        lo[0] = row_begin;
        rc = upload(ctx, &m->push_bits, bits.data(), bits.size(), bits.size(), &bytes);
        if (!rc) rc = upload(ctx, &m->push_lo, lo.data(), lo.size(), lo.size(), &bytes);
        if (!rc) rc = upload<uint32_t>(ctx, &m->push_count, nullptr, 0, n_blocks + 1, &bytes);
        m->n_push_blocks = n_blocks;
        // pusher CTAs (xchg_pusher_kernel): chunk-crossing rows whose chunks lie inside ONE push block are finished by the
        // CTA that sends the block (fix_in, grouped by block); the ones across block boundaries and the hub rows follow at
        // the end of the step (fix_def, fix_long: their bits are clear in pusher_bits); empty rows go with their block
        const uint32_t chunks_per_block = kPushCtas * kWarpsPerBlock;
        std::vector<glb_fixup_t> fix_in, fix_def;
        std::vector<uint32_t> blk_fs(n_blocks + 1, 0), blk_em(n_blocks + 1, 0);
        std::vector<uint32_t> pbits((size_t(row_end - base) + 31) / 32 + 1, 0u);
        for (uint32_t r = row_begin; r < row_end; ++r) pbits[(r - base) >> 5] |= 1u << ((r - base) & 31u);
        for (const auto &e : L.fix_short) {
            const uint32_t b0 = e.c_begin / chunks_per_block, b1 = (e.c_end & ~GLB_FLAG) / chunks_per_block;
            if (b0 == b1) {
                fix_in.push_back(e);
                blk_fs[b0 + 1]++;
            } else {
                fix_def.push_back(e);
                pbits[(e.row - base) >> 5] &= ~(1u << ((e.row - base) & 31u));
            }
        }
        for (const auto &e : L.fix_long) pbits[(e.row - base) >> 5] &= ~(1u << ((e.row - base) & 31u));
        for (uint32_t j = 0; j < n_blocks; ++j) blk_fs[j + 1] += blk_fs[j];
        for (uint32_t j = 0; j <= n_blocks; ++j)
            blk_em[j] = uint32_t(std::lower_bound(L.empty_rows.begin(), L.empty_rows.end(), lo[j]) - L.empty_rows.begin());
        blk_em[0] = 0;
        m->n_fix_def = uint32_t(fix_def.size());
        if (!rc) rc = upload(ctx, &m->fix_in, fix_in.data(), fix_in.size(), fix_in.size(), &bytes);
        if (!rc) rc = upload(ctx, &m->fix_def, fix_def.data(), fix_def.size(), fix_def.size(), &bytes);
        if (!rc) rc = upload(ctx, &m->blk_fs, blk_fs.data(), blk_fs.size(), blk_fs.size(), &bytes);
        if (!rc) rc = upload(ctx, &m->blk_em, blk_em.data(), blk_em.size(), blk_em.size(), &bytes);
        if (!rc) rc = upload(ctx, &m->pusher_bits, pbits.data(), pbits.size(), pbits.size(), &bytes);
    }
    {
        // sub-blocks of a split step (GLB_XCHG_SPLIT=<1..8>, default 4): chunk boundaries on CTA multiples, the row
        // open at each boundary, and where the (row-ordered) fix-up / empty-row lists cross those rows
        uint32_t S = env_u32("GLB_XCHG_SPLIT", 4);
        S = S < 1 ? 1 : S > GLB_MAX_SPLIT ? GLB_MAX_SPLIT : S;
        m->sub_chunk.assign(S + 1, L.n_chunks);
        m->sub_row.assign(S + 1, row_end);
        for (uint32_t sb = 0; sb < S; ++sb) {
            uint32_t c = uint32_t(uint64_t(L.n_chunks) * sb / S) / kWarpsPerBlock * kWarpsPerBlock;
            m->sub_chunk[sb] = c;
            const uint32_t k = c < L.n_chunks ? (L.chunk_first[c] & ~GLB_FLAG) : uint32_t(L.nz_rows.size());
            m->sub_row[sb] = sb == 0 ? row_begin : (k < L.nz_rows.size() ? L.nz_rows[k] : row_end);
        }
        auto cross = [&](const std::vector<uint32_t> &rows_of, std::vector<uint32_t> &out) {
            out.assign(S + 1, uint32_t(rows_of.size()));
            for (uint32_t sb = 0; sb <= S; ++sb)
                out[sb] = uint32_t(std::lower_bound(rows_of.begin(), rows_of.end(), m->sub_row[sb]) - rows_of.begin());
        };
        std::vector<uint32_t> rs, rl;
        for (const auto &e : L.fix_short) rs.push_back(e.row);
        for (const auto &e : L.fix_long) rl.push_back(e.row);
        cross(rs, m->sub_fs);
        cross(rl, m->sub_fl);
        cross(L.empty_rows, m->sub_em);
        m->sub_fs[0] = m->sub_fl[0] = m->sub_em[0] = 0;
    }
    if (!rc) rc = upload<float>(ctx, &m->head_carry, nullptr, 0, L.n_chunks, &bytes);
    if (!rc) rc = upload<float>(ctx, &m->tail_carry, nullptr, 0, L.n_chunks, &bytes);
    if (!rc && cudaStreamSynchronize(ctx->stream) != cudaSuccess) {
        glb_set_error("glb_csr_create: upload failed: %s", cudaGetErrorString(cudaGetLastError()));
        rc = GLB_ECUDA;
    }
    m->device_bytes = bytes;
    if (rc) { glb_csr_destroy(m); return rc; }
    *out = m;
    return GLB_OK;
}

int glb_csr_destroy(glb_csr_t m) {
    if (!m) return GLB_OK;
    cudaSetDevice(m->ctx->device);
    cudaStreamSynchronize(m->ctx->stream);
    cudaFree(m->stream); cudaFree(m->flags); cudaFree(m->chunk_goff); cudaFree(m->chunk_first); cudaFree(m->nz_rows);
    cudaFree(m->fix_short); cudaFree(m->fix_long); cudaFree(m->empty_rows);
    cudaFree(m->head_carry); cudaFree(m->tail_carry); cudaFree(m->hot_cols); cudaFree(m->hot_x); cudaFree(m->xbits); cudaFree(m->mbits);
    cudaFree(m->push_bits); cudaFree(m->push_lo); cudaFree(m->push_count);
    cudaFree(m->fix_in); cudaFree(m->fix_def); cudaFree(m->blk_fs); cudaFree(m->blk_em); cudaFree(m->pusher_bits);
    cudaFree(m->dx); cudaFree(m->dmask); cudaFree(m->dy);
    cudaFree(m->dx2); cudaFree(m->dmask2); cudaFree(m->dy2);
    glb_ctx_release(m->ctx);
    delete m;
    return GLB_OK;
}

int glb_csr_info(glb_csr_t m, uint64_t info[8]) {
    GLB_REQUIRE(m && info, "NULL argument");
    info[0] = m->row_end - m->row_begin;
    info[1] = m->num_cols;
    info[2] = m->nnz;
    info[3] = m->n_chunks;
    info[4] = uint64_t(m->n_fix_short) + m->n_fix_long;
    info[5] = m->n_empty;
    info[6] = m->device_bytes;
    info[7] = m->tile_k;
    return GLB_OK;
}

static int check_spmv_args(glb_ctx_t ctx, glb_csr_t m, int op, int mask_type, const float *x, const float *mask,
                           float *y, const glb_spmv_epilogue_t *ep) {
    GLB_REQUIRE(ctx && m && x && y, "NULL argument");
    GLB_REQUIRE(m->ctx == ctx, "matrix belongs to another context");
    GLB_REQUIRE(op >= GLB_OP_MUL_ADD && op <= GLB_OP_ADD_MIN, "invalid semiring op");
    GLB_REQUIRE(mask_type >= GLB_MASK_NONE && mask_type <= GLB_MASK_WRITE_TO_ONE, "invalid mask type");
    GLB_REQUIRE(mask_type == GLB_MASK_NONE || mask, "mask is NULL but mask_type != kNoMask");
    GLB_REQUIRE(x != y, "y must not alias x");
    if (ep && ep->assign_inout)
        GLB_REQUIRE(ep->assign_mask_type == GLB_MASK_WRITE_TO_ZERO || ep->assign_mask_type == GLB_MASK_WRITE_TO_ONE,
                    "fused dense assign needs a mask type");
    return GLB_OK;
}

int glb_spmv_fused(glb_ctx_t ctx, glb_csr_t m, int op, float zero, int mask_type, const float *x, const float *mask,
                   float *y, const glb_spmv_epilogue_t *ep) {
    int rc = check_spmv_args(ctx, m, op, mask_type, x, mask, y, ep);
    if (rc) return rc;
    return glb_launch_spmv(ctx, m, op, zero, mask_type, x, mask, y, ep, nullptr, 0, nullptr, nullptr, nullptr, GLB_VAL_F32, nullptr);
}

int glb_spmv_vt(glb_ctx_t ctx, glb_csr_t m, int val_type, int op, uint32_t zero_bits, int mask_type, const void *x, const void *mask,
                void *y, const glb_spmv_epilogue_t *ep) {
    GLB_REQUIRE(val_type >= GLB_VAL_F32 && val_type <= GLB_VAL_UFIXED, "invalid value type");
    float zero;
    memcpy(&zero, &zero_bits, sizeof(zero));
    int rc = check_spmv_args(ctx, m, op, mask_type, static_cast<const float *>(x), static_cast<const float *>(mask),
                             static_cast<float *>(y), ep);
    if (rc) return rc;
    return glb_launch_spmv(ctx, m, op, zero, mask_type, static_cast<const float *>(x), static_cast<const float *>(mask),
                           static_cast<float *>(y), ep, nullptr, 0, nullptr, nullptr, nullptr, val_type, nullptr);
}

int glb_spmv_exchange(glb_ctx_t ctx, glb_csr_t m, int op, float zero, int mask_type, glb_xchg_t xc, int src_vec, int dst_vec,
                      const float *mask, const glb_spmv_epilogue_t *ep) {
    return glb_spmv_exchange_iterate(ctx, m, op, zero, mask_type, xc, src_vec, dst_vec, mask, ep, 1, nullptr);
}

// n_steps row-sharded iterations x -> y -> x ... over two exchange vectors.  Per step: the SpMV
// kernels, whose first kernel opens with the acquire of the previous step's exchange, then the
// slice goes out:
//   multicast mapping:  one push kernel (16-byte multimem.st of the finished slice; its last CTA
//                       publishes the epoch).  GLB_XCHG_MC=fused stores each row to the multicast
//                       address from the SpMV write-back instead (measured slower: 4-byte remote stores);
//   peer mapping:       the write-back stored every row into all peers' blocks; one 32-thread kernel
//                       publishes the epoch.
// Only the last step is followed by an acquire of its own (so that whatever comes next on the
// stream sees the complete vector).  All launches are graph-recordable.
int glb_spmv_exchange_iterate(glb_ctx_t ctx, glb_csr_t m, int op, float zero, int mask_type, glb_xchg_t xc, int src_vec,
                              int dst_vec, const float *mask, const glb_spmv_epilogue_t *eps, int n_steps, const int *vec_plan) {
    GLB_REQUIRE(xc && xc->connected, "exchange is not connected");
    GLB_REQUIRE(src_vec >= 0 && src_vec < xc->n_vectors && dst_vec >= 0 && dst_vec < xc->n_vectors && src_vec != dst_vec,
                "bad vector index");
    for (int k = 0; vec_plan && k < 2 * n_steps; ++k)
        GLB_REQUIRE(vec_plan[k] >= 0 && vec_plan[k] < xc->n_vectors, "bad vector index in the plan");
    GLB_REQUIRE(m && m->num_rows <= xc->n && m->num_cols <= xc->n, "matrix larger than the exchange vectors");
    GLB_REQUIRE(n_steps >= 0, "negative step count");
    static const int forced = [] {
        const char *v = getenv("GLB_XCHG_MC");
        return !v ? 1 : !strcmp(v, "kernel") ? 1 : !strcmp(v, "fused") ? 2 : !strcmp(v, "progressive") ? 3 : !strcmp(v, "copy") ? 4 : !strcmp(v, "pusher") ? 5 : 1;
    }();
    const GlbXchgWait wait = glb_xchg_wait_desc(xc);
    float *peers[GLB_MAX_PEERS];
    for (int k = 0; k < n_steps; ++k) {
        const int sv = vec_plan ? vec_plan[2 * k] : (k & 1) ? dst_vec : src_vec;
        const int dv = vec_plan ? vec_plan[2 * k + 1] : (k & 1) ? src_vec : dst_vec;
        const float *x = xc->local + size_t(sv) * xc->n;
        float *y = xc->local + size_t(dv) * xc->n;
        const glb_spmv_epilogue_t *ep = eps ? eps + k : nullptr;
        int rc = check_spmv_args(ctx, m, op, mask_type, x, mask, y, ep);
        if (rc) return rc;
        const GlbXchgWait *w = k > 0 ? &wait : nullptr;
        // (per-kernel timing and the tile variant take the plain sequence)
        int mode = (forced == 4 && (ctx->timing || m->tile_threads)) ? 1 : forced;
        for (int r = 0; mode == 4 && r < xc->nranks; ++r)
            if (!xc->peer[r]) mode = 1;  // (an exchange made by glb_xchg_mc_open maps no peer block one by one)
        if (mode == 5 && (ctx->timing || !xc->mc || !glb_pusher_applies(m, y, xc->mc + size_t(dv) * xc->n))) mode = 1;
        if (xc->nranks > 1 && mode == 4) {
            // GLB_XCHG_MC=copy: split step -- sub-blocks of the shard on streams of their own, their finished rows
            // copied to the peers by the copy engines while later sub-blocks compute, then one publishing kernel
            GlbSpmvSplit sp;
            sp.n_peers = 0;
            sp.last_by_caller = xc->mc != nullptr;
            for (int r = 0; r < xc->nranks; ++r)
                if (r != xc->rank) sp.y_peers[sp.n_peers++] = xc->peer[r] + size_t(dv) * xc->n;
            rc = glb_launch_spmv(ctx, m, op, zero, mask_type, x, mask, y, ep, nullptr, 0, nullptr, w, nullptr, GLB_VAL_F32, &sp);
            if (!rc && sp.last_by_caller) {   // the last sub-block's rows by the push kernel, whose last CTA publishes
                const uint32_t r0 = m->sub_row[m->sub_row.size() - 2];
                rc = glb_xchg_push(ctx, xc, dv, r0, size_t(m->row_end - r0));
            } else if (!rc) {
                rc = glb_xchg_signal(ctx, xc, false);
            }
        } else if (xc->mc && xc->nranks > 1 && mode == 5) {
            // GLB_XCHG_MC=pusher: a few CTAs on SMs of their own finish and send every completed block of rows while
            // the later blocks compute, and publish (xchg_pusher_kernel; no fix-up launch)
            rc = glb_launch_pusher(ctx, m, op, zero, mask_type, mask, y, ep, xc->mc + size_t(dv) * xc->n, xc->mc_flags, xc->d_state, xc->rank);
            GlbSpmvMc mc;
            memset(&mc, 0, sizeof(mc));
            mc.pusher = true;
            if (!rc) rc = glb_launch_spmv(ctx, m, op, zero, mask_type, x, mask, y, ep, nullptr, 0, &mc, w, nullptr, GLB_VAL_F32, nullptr);
            if (!rc) GLB_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->pusher_ev_done, 0));
        } else if (xc->mc && xc->nranks > 1 && mode == 1) {
            // default: one push kernel after the SpMV kernels (all SMs store the finished slice in 16-byte
            // multimem.st; its last CTA publishes)
            rc = glb_launch_spmv(ctx, m, op, zero, mask_type, x, mask, y, ep, nullptr, 0, nullptr, w, nullptr, GLB_VAL_F32, nullptr);
            if (!rc) rc = glb_xchg_push(ctx, xc, dv, m->row_begin, size_t(m->row_end - m->row_begin));
        } else if (xc->mc && xc->nranks > 1) {
            // GLB_XCHG_MC=progressive: completed blocks of rows are pushed from inside the main kernel while
            // later blocks compute, the fix-up kernel sends its own rows and publishes; =fused: every row is
            // stored to the multicast address by the write-back that produces it.  Both overlap the transfer
            // with the kernel and both measured SLOWER than the separate push kernel (2 GPUs, C2: main kernel
            // 242 us vs 185 us, fix-up 29 us vs 8 us): remote stores issued from an SM that also runs compute
            // CTAs hold up that SM's L1 miss path, which is what bounds the SpMV.
            GlbSpmvMc mc;
            mc.y_mc = xc->mc + size_t(dv) * xc->n;
            mc.progressive = mode == 3;
            mc.pub_flags_mc = xc->mc_flags;
            mc.pub_state = xc->d_state;
            mc.rank = xc->rank;
            bool published = false;
            rc = glb_launch_spmv(ctx, m, op, zero, mask_type, x, mask, y, ep, nullptr, 0, &mc, w, &published, GLB_VAL_F32, nullptr);
            if (!rc && !published) rc = glb_xchg_signal(ctx, xc, false);
        } else {
            int n_peers = 0;
            for (int r = 0; r < xc->nranks; ++r)
                if (r != xc->rank) peers[n_peers++] = xc->peer[r] + size_t(dv) * xc->n;
            rc = glb_launch_spmv(ctx, m, op, zero, mask_type, x, mask, y, ep, peers, n_peers, nullptr, w, nullptr, GLB_VAL_F32, nullptr);
            if (!rc) rc = glb_xchg_signal(ctx, xc, false);
        }
        if (rc) return rc;
    }
    return n_steps ? glb_xchg_wait(ctx, xc) : GLB_OK;
}

int glb_spmv(glb_ctx_t ctx, glb_csr_t m, int op, float zero, int mask_type, const float *x, const float *mask,
             float *y) {
    return glb_spmv_fused(ctx, m, op, zero, mask_type, x, mask, y, nullptr);
}

int glb_spmv_host(glb_ctx_t ctx, glb_csr_t m, int op, float zero, int mask_type, const float *x_host,
                  const float *mask_host, float *y_host) {
    GLB_REQUIRE(ctx && m && x_host && y_host, "NULL argument");
    GLB_REQUIRE(m->ctx == ctx, "matrix belongs to another context");
    GLB_REQUIRE(mask_type == GLB_MASK_NONE || mask_host, "mask is NULL but mask_type != kNoMask");
    GLB_CUDA(cudaSetDevice(ctx->device));
    if (!m->dx) GLB_CUDA(cudaMalloc(reinterpret_cast<void **>(&m->dx), sizeof(float) * m->num_cols));
    if (mask_type != GLB_MASK_NONE && !m->dmask)
        GLB_CUDA(cudaMalloc(reinterpret_cast<void **>(&m->dmask), sizeof(float) * m->num_rows));
    // When y_host is page-locked (glb_host_alloc, cudaHostAlloc, torch pin_memory) the kernels store
    // the rows straight into it over PCIe (coalesced 128-byte writes), so the device-to-host leg
    // overlaps the SpMV instead of following it.
    float *y_dev = nullptr;
    cudaPointerAttributes attr;
    if (cudaPointerGetAttributes(&attr, y_host) == cudaSuccess && attr.type == cudaMemoryTypeHost && attr.devicePointer)
        y_dev = static_cast<float *>(attr.devicePointer);
    else
        cudaGetLastError();
    const bool zero_copy = y_dev != nullptr;
    if (!zero_copy) {
        if (!m->dy) GLB_CUDA(cudaMalloc(reinterpret_cast<void **>(&m->dy), sizeof(float) * m->num_rows));
        y_dev = m->dy;
    }
    GLB_CUDA(cudaMemcpyAsync(m->dx, x_host, sizeof(float) * m->num_cols, cudaMemcpyHostToDevice, ctx->stream));
    if (mask_type != GLB_MASK_NONE)
        GLB_CUDA(cudaMemcpyAsync(m->dmask, mask_host, sizeof(float) * m->num_rows, cudaMemcpyHostToDevice, ctx->stream));
    int rc = glb_spmv(ctx, m, op, zero, mask_type, m->dx, m->dmask, y_dev);
    if (rc) return rc;
    if (!zero_copy) {
        const size_t nr = size_t(m->row_end - m->row_begin);
        GLB_CUDA(cudaMemcpyAsync(y_host + m->row_begin, m->dy + m->row_begin, sizeof(float) * nr, cudaMemcpyDeviceToHost,
                                 ctx->stream));
    }
    GLB_CUDA(cudaStreamSynchronize(ctx->stream));
    return GLB_OK;
}

// Pipelined flavour of glb_spmv_host for a sequence of independent vectors: three streams (upload,
// kernels, download) over two device slots, so vector k+1 rides up the PCIe link and result k-1
// rides down while the kernels work on vector k.  Every vector is still copied host -> device and
// every result device -> host; what changes is that the three legs of consecutive vectors overlap
// (PCIe is full duplex and the copy engines run beside the SMs).  Host buffers should be page-locked
// (glb_host_alloc); pageable ones work but their copies do not overlap.
//
// With an exchange (row-sharded run, one process per GPU) the two slots are exchange vectors 0 / 1:
// every rank uploads only ITS slice of x over PCIe and the slices meet over NVLink
// (glb_xchg_allgather on the kernel stream), so the host link carries each x once per job instead of
// once per GPU.
static int spmv_host_batch(glb_ctx_t ctx, glb_csr_t m, int op, float zero, int mask_type, glb_xchg_t xc, int n_vectors,
                           const float *const *x_hosts, const float *const *mask_hosts, float *const *y_hosts) {
    GLB_REQUIRE(ctx && m && x_hosts && y_hosts && n_vectors >= 0, "bad argument");
    GLB_REQUIRE(m->ctx == ctx, "matrix belongs to another context");
    GLB_REQUIRE(mask_type == GLB_MASK_NONE || mask_hosts, "mask_hosts is NULL but mask_type != kNoMask");
    for (int k = 0; k < n_vectors; ++k) {
        GLB_REQUIRE(x_hosts[k] && y_hosts[k], "NULL vector in the batch");
        GLB_REQUIRE(mask_type == GLB_MASK_NONE || mask_hosts[k], "NULL mask in the batch");
    }
    size_t x_off = 0, x_cnt = m->num_cols;
    if (xc) {
        GLB_REQUIRE(xc->connected && xc->ctx == ctx && xc->n_vectors >= 2, "exchange is not connected to this context");
        GLB_REQUIRE(m->num_cols <= xc->n && xc->n % uint32_t(xc->nranks) == 0, "exchange vectors must divide evenly over the ranks");
        x_cnt = xc->n / uint32_t(xc->nranks);
        x_off = x_cnt * size_t(xc->rank);
    }
    if (n_vectors == 0) return GLB_OK;
    GLB_CUDA(cudaSetDevice(ctx->device));
    if (!ctx->copy_in) GLB_CUDA(cudaStreamCreateWithFlags(&ctx->copy_in, cudaStreamNonBlocking));
    if (!ctx->copy_out) GLB_CUDA(cudaStreamCreateWithFlags(&ctx->copy_out, cudaStreamNonBlocking));
    for (auto &row : ctx->pipe_ev)
        for (cudaEvent_t &e : row)
            if (!e) GLB_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    const bool masked = mask_type != GLB_MASK_NONE;
    float **dxs[2] = {&m->dx, &m->dx2}, **dms[2] = {&m->dmask, &m->dmask2}, **dys[2] = {&m->dy, &m->dy2};
    float *xslot[2];
    for (int s = 0; s < 2; ++s) {
        if (xc) {
            xslot[s] = xc->local + size_t(s) * xc->n;
        } else {
            if (!*dxs[s]) GLB_CUDA(cudaMalloc(reinterpret_cast<void **>(dxs[s]), sizeof(float) * m->num_cols));
            xslot[s] = *dxs[s];
        }
        if (!*dys[s]) GLB_CUDA(cudaMalloc(reinterpret_cast<void **>(dys[s]), sizeof(float) * m->num_rows));
        if (masked && !*dms[s]) GLB_CUDA(cudaMalloc(reinterpret_cast<void **>(dms[s]), sizeof(float) * m->num_rows));
    }
    cudaEvent_t *uploaded = ctx->pipe_ev[0], *computed = ctx->pipe_ev[1], *downloaded = ctx->pipe_ev[2];
    const size_t nr = size_t(m->row_end - m->row_begin);
    // work already queued on the caller's stream (an earlier glb_spmv_host may still read slot 0) goes first
    GLB_CUDA(cudaEventRecord(computed[0], ctx->stream));
    GLB_CUDA(cudaStreamWaitEvent(ctx->copy_in, computed[0], 0));
    if (xc) {  // no peer may still be reading what this batch overwrites
        int rc = glb_xchg_signal_wait(ctx, xc);
        if (rc) return rc;
    }
    // GLB_BATCH_TRACE=1: device timeline of the first vectors of the batch (upload / kernels / download), to stderr
    static const bool trace = getenv("GLB_BATCH_TRACE") != nullptr;
    // Over an exchange the copies of the neighbouring vectors (upload k+1, download k-1) are held back until the slice
    // push of vector k has left: the SM-issued NVLink stores of the push and the copy engines' PCIe traffic share the
    // GPU's path to the outside, and a push that takes 20 us alone took 130-200 us beside two saturated PCIe directions
    // (2 GPUs, C2; profiles/r2_host_batch_timeline_2gpu.txt).  GLB_BATCH_UNGATED=1 restores the free-running order.
    static const bool gated = getenv("GLB_BATCH_UNGATED") == nullptr;
    constexpr int kTraceVectors = 10;
    cudaEvent_t tev[1 + 7 * kTraceVectors] = {};
    auto stamp = [&](int k, int what, cudaStream_t st) {
        if (!trace || k >= kTraceVectors) return;
        cudaEvent_t &e = tev[k < 0 ? 0 : 1 + 7 * k + what];
        cudaEventCreate(&e);
        cudaEventRecord(e, st);
    };
    if (!ctx->pipe_pushed) GLB_CUDA(cudaEventCreateWithFlags(&ctx->pipe_pushed, cudaEventDisableTiming));
    stamp(-1, 0, ctx->stream);
    // upload k: slot k & 1 is free once the kernels of vector k-2 have read it
    auto enqueue_upload = [&](int k, bool after_push) -> int {
        const int s = k & 1;
        if (k >= 2) GLB_CUDA(cudaStreamWaitEvent(ctx->copy_in, computed[s], 0));
        if (after_push) GLB_CUDA(cudaStreamWaitEvent(ctx->copy_in, ctx->pipe_pushed, 0));
        stamp(k, 0, ctx->copy_in);
        if (x_off < m->num_cols) {
            const size_t cnt = (x_off + x_cnt <= m->num_cols) ? x_cnt : m->num_cols - x_off;
            GLB_CUDA(cudaMemcpyAsync(xslot[s] + x_off, x_hosts[k] + x_off, sizeof(float) * cnt, cudaMemcpyHostToDevice, ctx->copy_in));
        }
        stamp(k, 1, ctx->copy_in);
        if (masked)   // the mask is row-local: only the shard's rows are read
            GLB_CUDA(cudaMemcpyAsync(*dms[s] + m->row_begin, mask_hosts[k] + m->row_begin, sizeof(float) * nr,
                                     cudaMemcpyHostToDevice, ctx->copy_in));
        GLB_CUDA(cudaEventRecord(uploaded[s], ctx->copy_in));
        return GLB_OK;
    };
    auto enqueue_download = [&](int k, bool after_push) -> int {
        const int s = k & 1;
        GLB_CUDA(cudaStreamWaitEvent(ctx->copy_out, computed[s], 0));
        if (after_push) GLB_CUDA(cudaStreamWaitEvent(ctx->copy_out, ctx->pipe_pushed, 0));
        stamp(k, 5, ctx->copy_out);
        GLB_CUDA(cudaMemcpyAsync(y_hosts[k] + m->row_begin, *dys[s] + m->row_begin, sizeof(float) * nr, cudaMemcpyDeviceToHost,
                                 ctx->copy_out));
        stamp(k, 6, ctx->copy_out);
        GLB_CUDA(cudaEventRecord(downloaded[s], ctx->copy_out));
        return GLB_OK;
    };
    const bool hold = xc != nullptr && gated;
    int rc = enqueue_upload(0, false);
    if (rc) return rc;
    for (int k = 0; k < n_vectors; ++k) {
        const int s = k & 1;
        if (!hold && k + 1 < n_vectors && (rc = enqueue_upload(k + 1, false))) return rc;
        // kernels k: need the upload, and the download of result k-2 must have drained dy[s]
        GLB_CUDA(cudaStreamWaitEvent(ctx->stream, uploaded[s], 0));
        if (k >= 2) GLB_CUDA(cudaStreamWaitEvent(ctx->stream, downloaded[s], 0));
        stamp(k, 2, ctx->stream);
        if (xc) {  // the slices of x meet over NVLink; the wait also orders slot reuse across ranks
            rc = glb_xchg_push(ctx, xc, s, x_off, x_cnt);
            if (rc) return rc;
            GLB_CUDA(cudaEventRecord(ctx->pipe_pushed, ctx->stream));
            if (hold) {  // the copies that overlap this vector's SpMV start once its slice is out
                if (k + 1 < n_vectors && (rc = enqueue_upload(k + 1, true))) return rc;
                if (k >= 1 && (rc = enqueue_download(k - 1, true))) return rc;
            }
            rc = glb_xchg_wait(ctx, xc);
            if (rc) return rc;
        }
        stamp(k, 3, ctx->stream);
        rc = glb_spmv(ctx, m, op, zero, mask_type, xslot[s], masked ? *dms[s] : nullptr, *dys[s]);
        if (rc) return rc;
        stamp(k, 4, ctx->stream);
        GLB_CUDA(cudaEventRecord(computed[s], ctx->stream));
        if (!hold && (rc = enqueue_download(k, false))) return rc;
    }
    if (hold && (rc = enqueue_download(n_vectors - 1, false))) return rc;
    // the caller's stream ends after the last download, so stream-ordered work and events that follow see it
    GLB_CUDA(cudaStreamWaitEvent(ctx->stream, downloaded[(n_vectors - 1) & 1], 0));
    if (n_vectors >= 2) GLB_CUDA(cudaStreamWaitEvent(ctx->stream, downloaded[n_vectors & 1], 0));
    GLB_CUDA(cudaStreamSynchronize(ctx->stream));
    if (trace) {
        GLB_CUDA(cudaStreamSynchronize(ctx->copy_in));
        GLB_CUDA(cudaStreamSynchronize(ctx->copy_out));
        auto at = [&](int k, int what) {
            float ms = 0;
            cudaEvent_t e = tev[1 + 7 * k + what];
            if (!e || cudaEventElapsedTime(&ms, tev[0], e) != cudaSuccess) { cudaGetLastError(); return -1.0; }
            return double(ms) * 1e3;
        };
        for (int k = 0; k < n_vectors && k < kTraceVectors; ++k)
            fprintf(stderr, "[glb batch rank %d] vector %d: upload %.0f..%.0f us | exchange %.0f..%.0f | spmv ..%.0f | download %.0f..%.0f\n",
                    xc ? xc->rank : 0, k, at(k, 0), at(k, 1), at(k, 2), at(k, 3), at(k, 4), at(k, 5), at(k, 6));
        for (auto e : tev) if (e) cudaEventDestroy(e);
    }
    return GLB_OK;
}

int glb_spmv_host_batch(glb_ctx_t ctx, glb_csr_t m, int op, float zero, int mask_type, int n_vectors,
                        const float *const *x_hosts, const float *const *mask_hosts, float *const *y_hosts) {
    return spmv_host_batch(ctx, m, op, zero, mask_type, nullptr, n_vectors, x_hosts, mask_hosts, y_hosts);
}

int glb_spmv_host_batch_exchange(glb_ctx_t ctx, glb_csr_t m, int op, float zero, int mask_type, glb_xchg_t xc, int n_vectors,
                                 const float *const *x_hosts, const float *const *mask_hosts, float *const *y_hosts) {
    GLB_REQUIRE(xc, "exchange is NULL");
    return spmv_host_batch(ctx, m, op, zero, mask_type, xc, n_vectors, x_hosts, mask_hosts, y_hosts);
}

}  // extern "C"
