#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python tools/bench_apps.py bfs > gpurun_out/bench_apps_bfs.jsonl 2> gpurun_out/bench_apps_bfs.err; echo "apps rc=$?"
cut -c1-1100 gpurun_out/bench_apps_bfs.jsonl
mkdir -p /tmp/ds && python tools/make_dataset.py c3 /tmp/ds/c3.npz 2>&1 | tail -1
( echo "== bench_bfs c3 7"; timeout 300 benchmark/bin/bench_bfs 16 1024000 256000 30720 overlay.xclbin /tmp/ds/c3.npz 7 2>&1 | tail -8
  echo "== bench_sssp c3 6"; timeout 300 benchmark/bin/bench_sssp /tmp/ds/c3.npz 6 2>&1 | tail -8
  echo "== bench_pagerank c3"; timeout 300 benchmark/bin/bench_pagerank /tmp/ds/c3.npz 2>&1 | tail -5 ) > gpurun_out/cpp_bench.txt 2>&1
cat gpurun_out/cpp_bench.txt
