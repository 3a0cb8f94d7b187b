#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_siblings.py tests/test_gpu_step.py -m gpu -q -x -k "aug or color or step or flow_visualize" > gpurun_out/r02_pytest_sync.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r02_pytest_sync.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-kernel-rooflines --no-gpu-eager-baseline --timeline-out gpurun_out/r02_timeline_host_g1.txt > gpurun_out/r02_bench_sync.log 2>&1; echo "bench rc=$?"
python - <<'PY'
import json
for l in open('gpurun_out/r02_bench_sync.log'):
    if l.startswith('{'):
        d=json.loads(l); print("value %.1f clips/s %.2f ms/step; e2e %.1f clips/s %.2f ms; wall %.2f ms" % (d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['wall_ms_per_step']))
PY
head -2 gpurun_out/r02_timeline_host_g1.txt | cut -c1-220; grep -A8 "host side" gpurun_out/r02_timeline_host_g1.txt | cut -c1-150
