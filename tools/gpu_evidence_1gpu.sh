#!/bin/bash
# Final 1-GPU evidence: gpu tests, smoke, bench line, reference arm, ncu launch list + --set full of the dominant kernel, C++ drivers.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; cut -c1-150 gpurun_out/bench.json; grep -o '"e2e.*' gpurun_out/bench.json | cut -c1-160; grep -o '"roofline.*' gpurun_out/bench.json | cut -c1-400
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2>/dev/null; echo "ref rc=$?"; cut -c1-200 gpurun_out/bench_ref.json
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'spmv|gather_hot|pack_bits|fill|xchg' -c 400 --csv --log-file gpurun_out/launches_bench.csv \
    python bench.py --steps 20 --warmup 3 > gpurun_out/ncu_bench.log 2>&1; echo "list rc=$?"
python tools/launch_summary.py gpurun_out/launches_bench.csv
ncu --set full --clock-control none --import-source on -k regex:spmv_lane_kernel -s 2 -c 1 -f -o gpurun_out/prof_v3 \
    python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_v3.log 2>&1; echo "lane full rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'spmv|gather_hot|pack_bits|fill|spmspv|assign' -c 300 --csv --log-file gpurun_out/launches_bfs.csv \
    python tests/bench_apps.py bfs --no-check --reps 1 > gpurun_out/ncu_bfs.log 2>&1; echo "bfs list rc=$?"
mkdir -p /tmp/ds && python tools/make_dataset.py c3 /tmp/ds/c3.npz 2>&1 | tail -1
( echo "== bench_bfs c3 7"; timeout 300 benchmark/bin/bench_bfs 16 1024000 256000 30720 overlay.xclbin /tmp/ds/c3.npz 7 2>&1 | tail -8
  echo "== bench_sssp c3 6"; timeout 300 benchmark/bin/bench_sssp /tmp/ds/c3.npz 6 2>&1 | tail -8
  echo "== bench_pagerank c3"; timeout 300 benchmark/bin/bench_pagerank /tmp/ds/c3.npz 2>&1 | tail -5
  echo "== bench_spmv c3"; timeout 300 benchmark/bin/bench_spmv 16 1024000 30720 overlay.xclbin /tmp/ds/c3.npz 2>&1 | tail -4 ) > gpurun_out/cpp_bench.txt 2>&1
cat gpurun_out/cpp_bench.txt
