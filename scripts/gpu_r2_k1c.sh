#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_step.py -m gpu -q -x -k "infonce or objective or step" > gpurun_out/r02_pytest_k1c.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/r02_pytest_k1c.log
timeout 600 python - <<'PY'
import torch
from mscl_b200 import kernel_bench as kb
dev = torch.device("cuda", 0)
pk, _ = kb.hbm_peak()
rows = kb.bench_k1("cfg2", 96, 65536, pk, dev) + kb.bench_k1("cfg2", 32, 65536, pk, dev) + kb.bench_k1_pair("cfg2", (96, 32), 65536, pk, dev) + kb.bench_k1("cfg3", 64, 1048576, pk, dev, iters=20)
for r in rows:
    if "slab form" in r["kernel"]: continue
    print(f"{r['kernel'][:86]:<86} {r['shape']:<34} {r['us']:7.1f} us {100*r['frac_hbm']:5.1f}%")
PY
MSCL_TIMELINE=1 python -m mscl_b200.build --force > /dev/null 2>&1; echo "build rc=$?"
timeout 120 python scripts/tc_timeline_fused.py 96 65536 2 2>&1 | tail -32 > gpurun_out/r02_k1c_timeline.txt; cat gpurun_out/r02_k1c_timeline.txt
