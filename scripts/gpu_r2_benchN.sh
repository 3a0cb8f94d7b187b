#!/bin/bash
# round 2: the bench line on all GPUs of the box (sharded queue; REPL=1 also replicated), short
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
for mode in "" ${REPL:+--no-shard}; do
  tag=g$N$( [ -n "$mode" ] && echo _replicated )
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + N)) \
      bench.py --gpus $N --steps ${STEPS:-20} --warmup ${WARMUP:-5} --no-cpu-baseline --no-kernel-rooflines $mode > gpurun_out/r02_bench_$tag.log 2>&1
  echo "bench $tag rc=$?"
  python - "$tag" <<'PY'
import json, sys
tag = sys.argv[1]
for l in open(f"gpurun_out/r02_bench_{tag}.log"):
    if l.startswith("{"):
        d = json.loads(l)
        print("%s: value %.1f clips/s %.2f ms/step  e2e %.1f clips/s %.2f ms  queue=%s loss=%.4f" % (
            tag, d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["config"].get("queue"), d["loss"]))
PY
  grep -v '^{' gpurun_out/r02_bench_$tag.log | grep -i "error\|Traceback\|strides do not match" -A5 | head -12
done
