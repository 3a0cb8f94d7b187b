#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q -rs -k "$N or state_dict" > gpurun_out/r02_pytest_multi_g$N.log 2>&1; echo "pytest multi rc=$?"
tail -8 gpurun_out/r02_pytest_multi_g$N.log
