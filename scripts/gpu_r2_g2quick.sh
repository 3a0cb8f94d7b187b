#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
g=$(nvidia-smi -L | wc -l)
for rep in 1 2; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $g --master-addr 127.0.0.1 --master-port $((29500 + g + rep)) \
    bench.py --gpus $g --steps 20 --warmup 5 --no-cpu-baseline --no-kernel-rooflines > gpurun_out/r02_bench_quick_g$g.log 2>&1
python - "$g" <<'PY'
import json, sys
g = sys.argv[1]
for l in open(f"gpurun_out/r02_bench_quick_g{g}.log"):
    if l.startswith("{"):
        d = json.loads(l)
        print("N=%s value %.1f clips/s %.2f ms/step  e2e %.1f clips/s %.2f ms" % (g, d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"]))
PY
done
grep -v '^{' gpurun_out/r02_bench_quick_g$g.log | grep -i "error\|Traceback" -A4 | head
