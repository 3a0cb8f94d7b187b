#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
MSCL_TIMELINE=1 python -m mscl_b200.build --force > /dev/null 2>&1; echo "build rc=$?"
for args in "96 65536 2"; do
  echo "##### M K FLUSH = $args"
  timeout 120 python scripts/tc_timeline_fused.py $args 2>&1 | tail -31
done > gpurun_out/r2h_timeline.txt 2>&1
cat gpurun_out/r2h_timeline.txt
