#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu.log
timeout 600 python tools/bench_apps.py bfs pagerank sssp > gpurun_out/bench_apps_1gpu.jsonl 2> gpurun_out/bench_apps_1gpu.err; echo "apps rc=$?"
cut -c1-1300 gpurun_out/bench_apps_1gpu.jsonl
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'spmv|gather_hot|pack_bits|fill|spmspv|assign' -c 300 --csv --log-file gpurun_out/launches_pr_sssp.csv \
    python tools/bench_apps.py pagerank sssp --no-check --reps 1 > gpurun_out/ncu_pr_sssp.log 2>&1; echo "list rc=$?"
python tools/launch_summary.py gpurun_out/launches_pr_sssp.csv
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'spmv|gather_hot|pack_bits|fill|spmspv|assign' -c 300 --csv --log-file gpurun_out/launches_bfs.csv \
    python tools/bench_apps.py bfs --no-check --reps 1 > gpurun_out/ncu_bfs.log 2>&1; echo "list rc=$?"
grep "spmv_lane_bits" gpurun_out/launches_bfs.csv | awk -F'","' '{print $NF}' | tr -d '"' | head -40 | tr '\n' ' '
