#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
MSCL_TIMELINE=1 python -m mscl_b200.build --force > /dev/null 2>&1
(echo "##### M=96 K=65536"; python scripts/tc_timeline.py 96 65536 1 | tail -28
echo "##### M=32 K=65536"; python scripts/tc_timeline.py 32 65536 1 | tail -28
echo "##### M=128 K=1Mi burst 1"; python scripts/tc_timeline.py 128 1048576 1 | tail -28 | head -3
echo "##### M=128 K=1Mi burst 50"; python scripts/tc_timeline.py 128 1048576 50 | tail -28 | head -3
echo "##### M=32 K=1Mi burst 50"; python scripts/tc_timeline.py 32 1048576 50 | tail -28 | head -3) > gpurun_out/timeline_r02.txt 2>&1
cat gpurun_out/timeline_r02.txt
