#!/bin/bash
# timeline of the single-launch K1 (debug build on the box only)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
MSCL_TIMELINE=1 python -m mscl_b200.build --force > /dev/null 2>&1; echo "build rc=$?"
for args in "96 65536 1" "96 65536 0" "32 65536 0" "96 1048576 0"; do
  echo "##### M K FLUSH = $args"
  timeout 120 python scripts/tc_timeline_fused.py $args 2>&1 | tail -22
done > gpurun_out/r2b_timeline.txt 2>&1
cat gpurun_out/r2b_timeline.txt
