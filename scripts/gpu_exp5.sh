#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_step.py -m gpu -x -q 2>&1 | tail -15
timeout 600 python -m mscl_b200.kernel_bench --configs cfg2,cfg4 2>&1 | tail -60
