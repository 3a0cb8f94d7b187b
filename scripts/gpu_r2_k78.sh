#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_siblings.py -m gpu -q -x -k "trilinear or flow_visualize or color_pipeline or augment" > gpurun_out/r02_pytest_k78.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/r02_pytest_k78.log
timeout 600 python - <<'PY'
import torch, json
from mscl_b200 import kernel_bench as kb
dev = torch.device("cuda", 0)
pk, _ = kb.hbm_peak()
for r in kb.bench_k789("cfg2", 32, pk, dev):
    print(f"{r['kernel']:<58} {r['shape']:<52} {r['us']:7.1f} us {100*r['frac_hbm']:5.1f}%")
PY
SEL='(infonce or fra or lmcl) and not large_queue and not 65536 and not workspace and not 16384 and not 32768'
timeout 900 compute-sanitizer --tool racecheck --print-limit 20 --error-exitcode 7 \
      python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "$SEL" -p no:cacheprovider > gpurun_out/r02_sanitizer_racecheck.log 2>&1
echo "compute-sanitizer racecheck rc=$?"
grep -E "RACECHECK SUMMARY|passed|failed|Error:" gpurun_out/r02_sanitizer_racecheck.log | tail -4
