#!/bin/bash
# N-GPU check: exchange test, bench line with the peer exchange and with NCCL, app benches both ways.
N=${1:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
if [ "$N" = "2" ]; then timeout 600 python -m pytest tests/test_gpu_multi.py -x -q > gpurun_out/pytest_gpu_multi.log 2>&1; echo "pytest multi rc=$?"; tail -15 gpurun_out/pytest_gpu_multi.log; fi
timeout 600 $TR bench.py --gpus $N --steps 50 --warmup 5 > gpurun_out/bench_${N}gpu_peer.json 2> gpurun_out/bench_${N}gpu_peer.err; echo "bench peer rc=$?"
grep '^{' gpurun_out/bench_${N}gpu_peer.json | cut -c1-200
GLB_EXCHANGE=nccl timeout 600 $TR bench.py --gpus $N --steps 50 --warmup 5 > gpurun_out/bench_${N}gpu_nccl.json 2> gpurun_out/bench_${N}gpu_nccl.err; echo "bench nccl rc=$?"
grep '^{' gpurun_out/bench_${N}gpu_nccl.json | cut -c1-200
timeout 900 $TR tests/bench_apps.py ${APPS:-bfs pagerank sssp} ${APPFLAGS} > gpurun_out/bench_apps_${N}gpu_peer.jsonl 2> gpurun_out/bench_apps_${N}gpu_peer.err; echo "apps peer rc=$?"
grep '^{' gpurun_out/bench_apps_${N}gpu_peer.jsonl | cut -c1-900
tail -3 gpurun_out/bench_apps_${N}gpu_peer.err
GLB_EXCHANGE=nccl timeout 900 $TR tests/bench_apps.py ${APPS:-bfs pagerank sssp} --no-check > gpurun_out/bench_apps_${N}gpu_nccl.jsonl 2> gpurun_out/bench_apps_${N}gpu_nccl.err; echo "apps nccl rc=$?"
grep '^{' gpurun_out/bench_apps_${N}gpu_nccl.jsonl | cut -c1-900
