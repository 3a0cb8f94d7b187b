#!/bin/bash
# One GPU-box pass for the round's evidence: parity tests, smoke, full bench line, every kernel alone at the shapes of
# BASELINE configs 2-5, ncu launch list + full captures.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
bash scripts/gpu_check.sh
cp gpurun_out/bench.log gpurun_out/bench_full.log
# the reference's op sequence per block: on the box's host cores and as eager PyTorch on the GPU
timeout 600 python scripts/ref_blocks.py --device cpu > gpurun_out/ref_blocks_cpu.jsonl 2> gpurun_out/ref_blocks.err; echo "ref_blocks cpu rc=$?"
timeout 300 python scripts/ref_blocks.py --device cuda > gpurun_out/ref_blocks_cuda.jsonl 2>> gpurun_out/ref_blocks.err; echo "ref_blocks cuda rc=$?"
bash scripts/gpu_profile.sh
