#!/bin/bash
# ncu evidence: launch lists (bench.py step, BFS app) + one --set full capture of each dominant kernel.
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_bfs.csv \
    python tools/bench_apps.py bfs --no-check --reps 1 > gpurun_out/ncu_bfs.log 2>&1; echo "bfs list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:spmv_lane_bits -s 2 -c 1 -f -o gpurun_out/prof_bits \
    python tools/bench_apps.py bfs --no-check --reps 1 > gpurun_out/ncu_bits.log 2>&1; echo "bits full rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_bench.csv \
    python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_bench.log 2>&1; echo "bench list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:spmv_lane_kernel -s 2 -c 1 -f -o gpurun_out/prof_v3 \
    python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_v3.log 2>&1; echo "lane full rc=$?"
ls -la gpurun_out/*.ncu-rep
