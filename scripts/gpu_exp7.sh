#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -12
python scripts/time_aug.py 2>&1 | tail -9
bash scripts/gpu_quick.sh 2>&1 | tail -18 | head -6
