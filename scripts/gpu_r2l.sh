#!/bin/bash
# round 2: where the last CTA's finalize spends its time (debug builds shipped as libmscl_b200_tl*.so)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
for v in "" B C; do
  lib=$PWD/mscl_b200/lib/libmscl_b200_tl$v.so
  [ -f $lib ] || continue
  echo "########## variant '$v' (''= as shipped, B = 4 statistic copies, C = slab stores do not wait for the ticket)"
  MSCL_LIB=$lib timeout 120 python scripts/tc_timeline_fused.py 96 65536 2 2>&1 | tail -38 | grep -E "iter|softmax done|O full|stats warp|ticket|last CTA|epilogue|exit|finalize"
done > gpurun_out/r2l_timeline.txt 2>&1
cat gpurun_out/r2l_timeline.txt
