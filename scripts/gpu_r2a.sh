#!/bin/bash
# round 2, first pass over the single-launch K1: parity of every infonce test, then the queue sweep
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "infonce" -x > gpurun_out/r2a_pytest.log 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/r2a_pytest.log
timeout 300 python scripts/k1_sweep.py --Ks 65536,1048576 --Ms 32,96,128 --out gpurun_out/r2a_k1_sweep.json > gpurun_out/r2a_k1_sweep.log 2>&1; echo "sweep rc=$?"
cat gpurun_out/r2a_k1_sweep.log | tail -12
