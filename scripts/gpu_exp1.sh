#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 600 python scripts/k1_sweep.py > gpurun_out/k1_sweep.log 2>&1; echo "sweep rc=$?"
timeout 600 python bench.py --steps 10 --warmup 4 --no-cpu-baseline --channels-last > gpurun_out/bench_cl.log 2>&1; echo "bench_cl rc=$?"
cat gpurun_out/k1_sweep.log
tail -c 1500 gpurun_out/bench_cl.log | head -c 800
