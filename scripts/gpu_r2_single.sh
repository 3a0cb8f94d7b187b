#!/bin/bash
# round 2 single-GPU pass: all parity tests, then compute-sanitizer (memcheck, racecheck) over the InfoNCE / FRA / LMCL tests
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r02_pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/r02_pytest_gpu.log
SEL='(infonce or fra or lmcl) and not large_queue and not 65536 and not workspace and not 16384 and not 32768'
for tool in memcheck racecheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 --error-exitcode 7 \
      python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "$SEL" -p no:cacheprovider > gpurun_out/r02_sanitizer_$tool.log 2>&1
  echo "compute-sanitizer $tool rc=$?"
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|Error:" gpurun_out/r02_sanitizer_$tool.log | tail -5
done
