#!/bin/bash
# 1-GPU check: parity tests, bench line, app benches, C++ SpMSpV sweep.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" 
tail -5 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 100 --warmup 10 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
tail -c 2500 gpurun_out/bench.json
timeout 600 python tools/bench_apps.py bfs pagerank sssp > gpurun_out/bench_apps_1gpu.jsonl 2> gpurun_out/bench_apps_1gpu.err; echo "apps rc=$?"
cat gpurun_out/bench_apps_1gpu.jsonl | cut -c1-900
GLB_SPMV_BITS=0 timeout 300 python tools/bench_apps.py bfs --no-check > gpurun_out/bench_bfs_nobits.jsonl 2>/dev/null
cut -c1-700 gpurun_out/bench_bfs_nobits.jsonl
mkdir -p /tmp/ds && python tools/make_dataset.py c3 /tmp/ds/c3.npz > /dev/null 2>&1
timeout 300 benchmark/bin/bench_spmspv hw x.xclbin /tmp/ds/c3.npz > gpurun_out/cpp_bench_spmspv.txt 2>&1; echo "spmspv rc=$?"
cat gpurun_out/cpp_bench_spmspv.txt | tail -12
