#!/bin/bash
# One GPU-box pass for the round's evidence: parity tests, smoke, full bench line, ncu launch list + full captures.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
bash scripts/gpu_check.sh
cp gpurun_out/bench.log gpurun_out/bench_full.log
bash scripts/gpu_profile.sh
