#!/bin/bash
# round 2 multi-GPU pass (gpurun --gpus 8): parity at world 2/4/8, the K1 sweep sharded vs replicated at 2/4/8, the bench at
# 8 with the communication timeline, and at 2 / 4.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
echo "GPUs: $N"
WL=${WORLD_LIST:-$N}
SEL=${PYTEST_K:-"$N or state_dict"}
timeout 1200 python -m pytest tests/test_gpu_multi.py -m gpu -q -rs -k "$SEL" > gpurun_out/r02_pytest_multi_g$N.log 2>&1; echo "pytest multi rc=$?"
tail -6 gpurun_out/r02_pytest_multi_g$N.log
for g in $WL; do
  [ "$g" -le "$N" ] || continue
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $g --master-addr 127.0.0.1 --master-port $((29600 + g)) \
      scripts/k1_sweep_multi.py --out gpurun_out/r02_k1_sweep_g$g.json > gpurun_out/r02_k1_sweep_g$g.log 2>&1
  echo "k1 sweep g=$g rc=$?"; grep "^G=" gpurun_out/r02_k1_sweep_g$g.log
done
for g in $WL; do
  [ "$g" -le "$N" ] || continue
  extra="--timeline-out gpurun_out/r02_timeline_g$g.txt"
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $g --master-addr 127.0.0.1 --master-port $((29500 + g)) \
      bench.py --gpus $g --steps 20 --warmup 5 --no-cpu-baseline --no-kernel-rooflines $extra > gpurun_out/r02_bench_g$g.log 2>&1
  echo "bench g=$g rc=$?"
  python - "$g" <<'PY'
import json, sys
g = sys.argv[1]
for l in open(f"gpurun_out/r02_bench_g{g}.log"):
    if l.startswith("{"):
        d = json.loads(l)
        print("N=%s value %.1f clips/s %.2f ms/step  e2e %.1f clips/s  queue=%s loss=%.4f" % (
            g, d["value"], d["ms_per_step"], d["e2e"]["value"], d["config"].get("queue"), d["loss"]))
PY
  grep -c "Grad strides do not match" gpurun_out/r02_bench_g$g.log
  grep -v '^{' gpurun_out/r02_bench_g$g.log | grep -i "error\|Traceback" | head -5
done
head -14 gpurun_out/r02_timeline_g$N.txt
