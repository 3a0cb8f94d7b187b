#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
python - <<'PY'
import torch
from mscl_b200 import kernel_bench as kb
pk,_=kb.hbm_peak()
dev=torch.device("cuda",0)
for r in kb.bench_k789("cfg2",32,pk,dev):
    print("%-34s %-50s %8.1f us %6.0f GB/s %5.1f%%"%(r["kernel"], r["shape"], r["us"], r["gbs"], 100*r["frac_hbm"]))
PY
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "flow_vis or color or augment" 2>&1 | tail -2
