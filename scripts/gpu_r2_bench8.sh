#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
g=$(nvidia-smi -L | wc -l)
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $g --master-addr 127.0.0.1 --master-port $((29500 + g)) \
    bench.py --gpus $g --steps 20 --warmup 5 --no-cpu-baseline --no-kernel-rooflines --timeline-out gpurun_out/r02_timeline_g$g.txt > gpurun_out/r02_bench_g$g.log 2>&1
echo "bench g=$g rc=$?"
python - "$g" <<'PY'
import json, sys
g = sys.argv[1]
for l in open(f"gpurun_out/r02_bench_g{g}.log"):
    if l.startswith("{"):
        d = json.loads(l)
        print("N=%s value %.1f clips/s %.2f ms/step  e2e %.1f clips/s %.2f ms  queue=%s loss=%.4f" % (
            g, d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["config"].get("queue"), d["loss"]))
PY
grep -c "Grad strides do not match" gpurun_out/r02_bench_g$g.log
grep -v '^{' gpurun_out/r02_bench_g$g.log | grep -i "error\|Traceback" -A4 | head -12
head -12 gpurun_out/r02_timeline_g$g.txt
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $g --master-addr 127.0.0.1 --master-port $((29700 + g)) \
    bench.py --gpus $g --steps 20 --warmup 5 --no-cpu-baseline --no-kernel-rooflines --no-shard > gpurun_out/r02_bench_g${g}_replicated.log 2>&1
python - "$g" <<'PY'
import json, sys
g = sys.argv[1]
for l in open(f"gpurun_out/r02_bench_g{g}_replicated.log"):
    if l.startswith("{"):
        d = json.loads(l)
        print("N=%s REPLICATED queue: value %.1f clips/s %.2f ms/step" % (g, d["value"], d["ms_per_step"]))
PY
