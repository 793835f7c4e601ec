#!/bin/bash
# Round-2 1-GPU evidence: gpu tests, smoke, bench line (SpMV + app records), reference arm, ncu launch lists,
# the C++ drivers of benchmark/ and the reference's own drivers compiled unmodified, SpMSpV sweeps.
mkdir -p gpurun_out /tmp/ds
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r2_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python bench.py > gpurun_out/r2_bench_1gpu.json 2> gpurun_out/r2_bench_1gpu.err; echo "bench rc=$?"; cut -c1-120 gpurun_out/r2_bench_1gpu.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_reference_arm.json 2>/dev/null; echo "ref rc=$?"; cut -c1-200 gpurun_out/r2_bench_reference_arm.json
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'spmv|gather_hot|pack_bits|fill|xchg|spmspv|assign|sparse' -c 400 --csv --log-file gpurun_out/r2_launches_bench.csv \
    python bench.py --steps 20 --warmup 3 --no-apps > gpurun_out/r2_ncu_bench.log 2>&1; echo "list rc=$?"
python tools/launch_summary.py gpurun_out/r2_launches_bench.csv > gpurun_out/r2_launches_bench_summary.txt; cat gpurun_out/r2_launches_bench_summary.txt
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'spmv|gather_hot|pack_bits|fill|spmspv|assign|sparse' -c 300 --csv --log-file gpurun_out/r2_launches_bfs_app.csv \
    python tests/bench_apps.py bfs --no-check --reps 1 > gpurun_out/r2_ncu_bfs.log 2>&1; echo "bfs list rc=$?"
python tools/launch_summary.py gpurun_out/r2_launches_bfs_app.csv > gpurun_out/r2_launches_bfs_app_summary.txt; cat gpurun_out/r2_launches_bfs_app_summary.txt
python tools/make_dataset.py c3 /tmp/ds/c3.npz 2>&1 | tail -1
python tools/make_dataset.py c5 /tmp/ds/c5.npz 0.25 2>&1 | tail -1
( echo "== benchmark/bin/bench_bfs c3 7"; timeout 300 benchmark/bin/bench_bfs 16 1024000 256000 30720 overlay.xclbin /tmp/ds/c3.npz 7 2>&1 | tail -8
  echo "== benchmark/bin/bench_sssp c3 6"; timeout 300 benchmark/bin/bench_sssp /tmp/ds/c3.npz 6 2>&1 | tail -8
  echo "== benchmark/bin/bench_pagerank c3"; timeout 300 benchmark/bin/bench_pagerank /tmp/ds/c3.npz 2>&1 | tail -5
  echo "== benchmark/bin/bench_spmv c3"; timeout 300 benchmark/bin/bench_spmv 16 1024000 30720 overlay.xclbin /tmp/ds/c3.npz 2>&1 | tail -4 ) > gpurun_out/r2_cpp_benchmark_drivers.txt 2>&1
cat gpurun_out/r2_cpp_benchmark_drivers.txt
# the reference's OWN drivers, compiled unmodified (tests/cpp/Makefile refbench_*), on the same dataset
( echo "== /root/reference/benchmark/bench_spmv.cpp (unmodified) c3"; timeout 300 tests/cpp/bin/refbench_bench_spmv 16 1024000 30720 overlay.xclbin /tmp/ds/c3.npz 2>&1 | tail -6
  echo "== /root/reference/benchmark/bench_bfs.cpp (unmodified) c3 7"; timeout 300 tests/cpp/bin/refbench_bench_bfs 16 1024000 256000 30720 overlay.xclbin /tmp/ds/c3.npz 7 2>&1 | tail -8
  echo "== /root/reference/benchmark/bench_pagerank.cpp (unmodified) c3"; timeout 300 tests/cpp/bin/refbench_bench_pagerank 16 1024000 30720 overlay.xclbin /tmp/ds/c3.npz 2>&1 | tail -6
  echo "== /root/reference/benchmark/bench_sssp.cpp (unmodified) c3 6"; timeout 300 tests/cpp/bin/refbench_bench_sssp 16 1024000 256000 30720 overlay.xclbin /tmp/ds/c3.npz 6 2>&1 | tail -8 ) > gpurun_out/r2_reference_drivers_unmodified.txt 2>&1
cat gpurun_out/r2_reference_drivers_unmodified.txt
( GLB_HBM_PEAK_GBS=6551 timeout 300 benchmark/bin/bench_spmspv /tmp/ds/c3.npz /tmp/ds/c5.npz ) > gpurun_out/r2_cpp_bench_spmspv.txt 2>&1; tail -60 gpurun_out/r2_cpp_bench_spmspv.txt
