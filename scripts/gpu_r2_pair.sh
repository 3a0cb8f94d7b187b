#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_step.py tests/test_gpu_siblings.py -m gpu -q -x -k "infonce or objective or step or mscl or modist" > gpurun_out/r02_pytest_pair.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/r02_pytest_pair.log
timeout 600 python - <<'PY'
import torch
from mscl_b200 import kernel_bench as kb
dev = torch.device("cuda", 0)
pk, _ = kb.hbm_peak()
for r in kb.bench_k1("cfg2", 96, 65536, pk, dev) + kb.bench_k1_pair("cfg2", (96, 32), 65536, pk, dev) + kb.bench_k1_pair("cfg2", (96, 96), 65536, pk, dev):
    print(f"{r['kernel'][:86]:<86} {r['shape']:<34} {r['us']:7.1f} us {100*r['frac_hbm']:5.1f}%")
PY
