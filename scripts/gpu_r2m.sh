#!/bin/bash
# round 2: where the GPU idles inside the N = 1 step (torch.profiler gap list), and the in-step K1 timings with the tensor-map cache
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_siblings.py tests/test_gpu_step.py -m gpu -q -x -k "color or augment or aug or flow_vis or train_step or parse" 2>&1 | tail -3
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-kernel-rooflines --no-gpu-eager-baseline --timeline-out gpurun_out/r02_timeline_host_g1.txt > gpurun_out/r02_bench_host_g1.log 2>&1; echo "bench rc=$?"
python - <<'PY'
import json
for l in open("gpurun_out/r02_bench_host_g1.log"):
    if l.startswith("{"):
        d = json.loads(l)
        print("value %.1f clips/s %.2f ms/step e2e %.1f" % (d["value"], d["ms_per_step"], d["e2e"]["value"]))
        print(json.dumps(d["roofline"])[:900])
        for k in d["kernels"]:
            if "infonce" in k["kernel"]: print(k["kernel"], round(k["avg_us"], 1))
PY
grep -v '^{' gpurun_out/r02_bench_host_g1.log | grep -i "error\|Traceback" -A5 | head
sed -n 1,3p gpurun_out/r02_timeline_host_g1.txt; grep -A60 "^# idle gaps" gpurun_out/r02_timeline_host_g1.txt | head -90
