#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "infonce" -x > gpurun_out/r2j_pytest.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/r2j_pytest.log
timeout 300 python scripts/k1_sweep.py --Ks 65536,1048576 --Ms 32,96,128 --out gpurun_out/r2j_k1_sweep.json > gpurun_out/r2j_k1_sweep.log 2>&1; echo "sweep rc=$?"
cat gpurun_out/r2j_k1_sweep.log | tail -8
MSCL_TIMELINE=1 python -m mscl_b200.build --force > /dev/null 2>&1; echo "build rc=$?"
for args in "96 65536 2" "96 65536 0"; do
  echo "##### M K FLUSH = $args"
  timeout 120 python scripts/tc_timeline_fused.py $args 2>&1 | tail -30
done > gpurun_out/r2j_timeline.txt 2>&1
cat gpurun_out/r2j_timeline.txt
