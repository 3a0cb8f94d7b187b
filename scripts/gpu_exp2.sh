#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
for v in BASE NOLOAD2 NOMMA2 NOSOFTMAX "NOLOAD2 -DMSCL_EXP_NOMMA2"; do
  if [ "$v" = BASE ]; then defs=""; else defs="-DMSCL_EXP_$v"; fi
  MSCL_DEFS="$defs" python -m mscl_b200.build --force > /dev/null 2>&1
  echo "=== variant $v"
  timeout 300 python scripts/k1_sweep.py --Ks 65536,1048576 --Ms 32,96,128 --out gpurun_out/k1_exp_${v// /_}.json 2>&1 | cut -c1-175
done
python -m mscl_b200.build --force > /dev/null 2>&1
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q 2>&1 | tail -3
