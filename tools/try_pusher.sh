#!/bin/bash
# Exchange-mode comparison at N ranks (gpurun --gpus N): the SpMV bench line only, parity checked by every rank.
#   bash tools/try_pusher.sh <N>            kernel (default exchange) vs pusher CTAs with 4 and 8 pushers
#   CFGS="3 6" GLB_XCHG_TRACE=6 bash ...    other pusher counts; timeline of the 6th eager step of rank 0
N=${1:-2}
mkdir -p gpurun_out
run() {  # run <tag> <env...>
    tag=$1; shift
    env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 \
        bench.py --gpus $N --steps 100 --warmup 10 --no-apps > gpurun_out/pusher_${N}gpu_$tag.json 2> gpurun_out/pusher_${N}gpu_$tag.err
    echo "$tag rc=$? $(grep '^{' gpurun_out/pusher_${N}gpu_$tag.json | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('GTEPS', round(d['value'],1), 'ms/step', round(d['ms_per_step'],4), 'parity', d['parity']['ok'], 'launched', round(d['step_breakdown']['host_launched_ms_per_step'],4), 'main+fix', round(d['step_breakdown']['main_plus_fixup_ms_max_over_ranks'],4), 'e2e', round(d['e2e']['value'],1))
" 2>&1)"
    grep -i "error\|trap\|illegal\|Traceback" gpurun_out/pusher_${N}gpu_$tag.err | head -3
    grep "glb pusher rank 0" gpurun_out/pusher_${N}gpu_$tag.err | head -40
}
[ -n "$SKIP_KERNEL" ] || run kernel GLB_XCHG_MC=kernel
CFGS=${CFGS:-4 8}
for cfg in $CFGS; do
    run pusher$cfg GLB_XCHG_MC=pusher GLB_XCHG_PUSHERS=$cfg
done
