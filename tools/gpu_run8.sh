#!/bin/bash
# 8-GPU (or N-GPU) run: bench line with the peer exchange and with NCCL, app benches over the peer exchange.
N=${1:-8}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 400 $TR bench.py --gpus $N --steps 50 --warmup 5 > gpurun_out/bench_${N}gpu_peer.json 2> gpurun_out/bench_${N}gpu_peer.err; echo "bench peer rc=$?"
grep '^{' gpurun_out/bench_${N}gpu_peer.json | cut -c1-200; grep -o '"e2e.*' gpurun_out/bench_${N}gpu_peer.json | cut -c1-300
GLB_EXCHANGE=nccl timeout 400 $TR bench.py --gpus $N --steps 50 --warmup 5 > gpurun_out/bench_${N}gpu_nccl.json 2> gpurun_out/bench_${N}gpu_nccl.err; echo "bench nccl rc=$?"
grep '^{' gpurun_out/bench_${N}gpu_nccl.json | cut -c1-200
timeout 600 $TR tools/bench_apps.py bfs pagerank sssp > gpurun_out/bench_apps_${N}gpu_peer.jsonl 2> gpurun_out/bench_apps_${N}gpu_peer.err; echo "apps peer rc=$?"
grep '^{' gpurun_out/bench_apps_${N}gpu_peer.jsonl | cut -c1-1000
tail -3 gpurun_out/bench_apps_${N}gpu_peer.err
