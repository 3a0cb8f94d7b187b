#!/bin/bash
# Multi-GPU pass (gpurun --gpus N): sharded-queue parity tests, then the bench at every power of two up to N.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
echo "GPUs: $N"
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q > gpurun_out/pytest_multi.log 2>&1; echo "pytest multi rc=$?"
tail -3 gpurun_out/pytest_multi.log
for g in ${GPU_LIST:-2 4 8}; do
  [ "$g" -le "$N" ] || continue
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $g --master-addr 127.0.0.1 --master-port $((29500 + g)) \
      bench.py --gpus $g --steps 10 --warmup 3 --no-cpu-baseline --no-kernel-rooflines > gpurun_out/bench_g$g.log 2>&1
  echo "bench g=$g rc=$?"
  python - "$g" <<'PY'
import json, sys
g = sys.argv[1]
for l in open(f"gpurun_out/bench_g{g}.log"):
    if l.startswith("{"):
        d = json.loads(l)
        print("N=%s value %.1f clips/s %.2f ms/step  e2e %.1f clips/s  queue=%s loss=%.4f" % (
            g, d["value"], d["ms_per_step"], d["e2e"]["value"], d["config"].get("queue"), d["loss"]))
PY
  grep -v '^{' gpurun_out/bench_g$g.log | grep -i "error\|Traceback" | head -5
done
