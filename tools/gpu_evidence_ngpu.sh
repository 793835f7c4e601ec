#!/bin/bash
# 8-GPU (or N-GPU) run: bench line over the multicast exchange (fused / separate kernel), app benches.
N=${1:-8}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
show() { grep '^{' $1 | python -c "
import sys, json
d=json.loads(sys.stdin.read()); r=d['roofline']; e=d['e2e']
print(d['n_gpus'], 'GTEPS', round(d['value'],1), 'ms/step', round(d['ms_per_step'],4), 'kernel', round(r['kernel_ms'],4), 'fixup', round(r['fixup_kernel_ms'],4), '| e2e', round(e['value'],1), round(e['ms_per_step'],4), '|', d['config'].get('exchange','')[:60])"; }
timeout 400 $TR bench.py --gpus $N --steps 50 --warmup 5 > gpurun_out/bench_${N}gpu_mc.json 2> gpurun_out/bench_${N}gpu_mc.err; echo "bench mc rc=$?"; show gpurun_out/bench_${N}gpu_mc.json
grep -i "unavailable" gpurun_out/bench_${N}gpu_mc.err | head -3
GLB_XCHG_MC=kernel timeout 400 $TR bench.py --gpus $N --steps 50 --warmup 5 > gpurun_out/bench_${N}gpu_mck.json 2> gpurun_out/bench_${N}gpu_mck.err; echo "bench mc-kernel rc=$?"; show gpurun_out/bench_${N}gpu_mck.json
timeout 600 $TR tests/bench_apps.py bfs pagerank sssp > gpurun_out/bench_apps_${N}gpu_mc.jsonl 2> gpurun_out/bench_apps_${N}gpu_mc.err; echo "apps rc=$?"
grep '^{' gpurun_out/bench_apps_${N}gpu_mc.jsonl | python -c "
import sys, json
for l in sys.stdin:
    d=json.loads(l); p=d['modes']['pull']
    print(d['app'], d['n_gpus'], 'whole ms/iter', round(p['ms_per_iteration'],4), 'loop ms/iter', round(p['loop_only_ms_per_iteration'],4), 'loop gteps', round(p['loop_only_gteps'],1), 'kernels', round(p['spmv_kernels_ms_per_iteration_max_over_ranks'],4), 'ok', (d['matches_oracle'] or {}).get('pull'), d['config']['sharding'][:50])"
tail -2 gpurun_out/bench_apps_${N}gpu_mc.err
