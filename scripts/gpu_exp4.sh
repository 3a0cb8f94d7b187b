#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_step.py -m gpu -x -q 2>&1 | tail -3
timeout 300 python scripts/k1_sweep.py --Ks 16384,65536,1048576 --Ms 32,96,128 --out gpurun_out/k1_pdl.json 2>&1 | cut -c1-175
