#!/bin/bash
# round 2: K9 after the indexing rework -- parity tests of the augmentation kernels + stand-alone timings
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_siblings.py -m gpu -q -x -k "color or augment or aug or flow_vis" 2>&1 | tail -3
timeout 300 python - <<'PY' 2>&1 | tee gpurun_out/r2n_k789.txt
import torch
from mscl_b200 import kernel_bench as kb
dev = torch.device("cuda", 0)
pk, _ = kb.hbm_peak()
for r in kb.bench_k789("cfg2", 32, pk, dev):
    if "color" in r["kernel"] or "flow_vis" in r["kernel"]:
        print(f"{r['kernel'][:66]:<66} {r['shape']:<52} {r['us']:7.1f} us {100*r['frac_hbm']:5.1f}%")
PY
