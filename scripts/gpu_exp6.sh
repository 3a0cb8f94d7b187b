#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
for cfgs in "1 256" "1 512" "1 1024" "2 512"; do
  set -- $cfgs
  echo "== cluster $1 threads $2"
  MSCL_FRA_CLUSTER=$1 MSCL_FRA_THREADS=$2 python - <<'PY' 2>&1 | grep fra_fused
import torch
from mscl_b200 import kernel_bench as kb
pk,_=kb.hbm_peak()
dev=torch.device("cuda",0)
for cfg,N,T in (("cfg2",32,8),("cfg4",64,16)):
    for r in kb.bench_k3(cfg,N,T,pk,dev):
        print(r["config"], r["kernel"], "%.1f us %.0f GB/s %.1f%%"%(r["us"], r["gbs"], 100*r["frac_hbm"]))
PY
done
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k fra 2>&1 | tail -2
