#!/bin/bash
# round 2, epoch-split K1: parity (kernel + objective + step + siblings), stand-alone timings, timelines (debug build shipped as libmscl_b200_tl.so)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_step.py tests/test_gpu_siblings.py -m gpu -q -x -k "infonce or objective or step or mscl or modist" > gpurun_out/r2k_pytest.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/r2k_pytest.log
timeout 600 python - <<'PY' 2>&1 | tee gpurun_out/r2k_k1.txt
import torch
from mscl_b200 import kernel_bench as kb
dev = torch.device("cuda", 0)
pk, _ = kb.hbm_peak()
rows = kb.bench_k1("cfg2", 96, 65536, pk, dev) + kb.bench_k1_pair("cfg2", (96, 32), 65536, pk, dev) + kb.bench_k1_step("cfg2", 32, 65536, pk, dev)
for r in rows:
    if "slab form" in r["kernel"]: continue
    print(f"{r['kernel'][:100]:<100} {r['shape']:<36} {r['us']:7.1f} us {100*r['frac_hbm']:5.1f}%  {r.get('note','')}")
PY
if [ -f mscl_b200/lib/libmscl_b200_tl.so ]; then
  export MSCL_LIB=$PWD/mscl_b200/lib/libmscl_b200_tl.so
  (echo "##### single op M=96 K=65536 (one launch after a device sync)"; timeout 120 python scripts/tc_timeline_fused.py 96 65536 2 2>&1 | tail -36
   echo "##### step launch: 96 rows over W_rgb + (32 | 96) rows over W_flow, K=65536"; timeout 120 python scripts/tc_timeline_fused.py 96 65536 2 step 2>&1 | tail -36) > gpurun_out/r2k_timeline.txt
  cat gpurun_out/r2k_timeline.txt
fi
