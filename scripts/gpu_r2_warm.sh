#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
if [ "$N" = 1 ]; then
  timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-kernel-rooflines --no-gpu-eager-baseline > gpurun_out/r02_bench_w3_g1.log 2>&1; echo "rc=$?"
else
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline --no-kernel-rooflines > gpurun_out/r02_bench_w3_g$N.log 2>&1; echo "rc=$?"
fi
python - "$N" <<'PY'
import json, sys
for l in open(f"gpurun_out/r02_bench_w3_g{sys.argv[1]}.log"):
    if l.startswith("{"):
        d = json.loads(l)
        print("N=%s W=3 K=10: value %.1f clips/s %.2f ms/step  e2e %.1f clips/s %.2f ms" % (sys.argv[1], d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"]))
PY
