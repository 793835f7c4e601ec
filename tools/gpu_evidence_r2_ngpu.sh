#!/bin/bash
# Round-2 N-GPU evidence (gpurun --gpus N): the bench line at N ranks (SpMV + app records, parity on every rank)
N=${1:-8}
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 \
    bench.py --gpus $N --steps 100 --warmup 10 > gpurun_out/r2_bench_${N}gpu.json 2> gpurun_out/r2_bench_${N}gpu.err
echo "bench rc=$?"; grep -v "^\[W\|Warning\|^\*\*\*" gpurun_out/r2_bench_${N}gpu.err | tail -6
grep "^{" gpurun_out/r2_bench_${N}gpu.json | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('value', d['value'], 'ms/step', d['ms_per_step'], 'parity', d['parity']['ok'], 'frac', d['roofline']['frac'], d['step_breakdown'])
print('e2e', d['e2e']['value'], d['e2e']['ms_per_step'])
for k in ('bfs','pagerank','sssp'):
    if k in d: print(k, d[k].get('parity'), {m:(round(v['iterations_per_sec']), round(v['ms_total'],4)) for m,v in d[k].get('modes',{}).items()})
"
# what the host side of the box sustains with N GPUs copying at once (the ceiling of the e2e figure)
timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29561 \
    tools/host_link_bw.py 2>/dev/null | grep "GPU(s)" | tee gpurun_out/r2_host_link_bw_${N}gpu.txt
if [ "$N" = "2" ]; then
  (timeout 400 python -m pytest tests/test_gpu_multi.py tests/test_cpp_host.py -q -k "two_gpus or sharded" 2>&1 | tail -5) | tee gpurun_out/r2_pytest_gpu_multi_2gpu.log
fi
