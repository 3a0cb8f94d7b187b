#!/bin/bash
# round 2, final multi-GPU pass on the N GPUs of the box: parity tests at world N (sharded and replicated queues, shuffle,
# one-rank state_dict), the bench with sharded and with replicated queues.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
timeout 1200 python -m pytest tests/test_gpu_multi.py -m gpu -q -rs -k "$N or state_dict" > gpurun_out/r02_pytest_multi_g$N.log 2>&1; echo "pytest multi rc=$?"
tail -5 gpurun_out/r02_pytest_multi_g$N.log
for mode in "" "--no-shard"; do
  tag=g$N$( [ -n "$mode" ] && echo _replicated )
  extra=$( [ -z "$mode" ] && echo "--timeline-out gpurun_out/r02_timeline_g$N.txt" )
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + N)) \
      bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline --no-kernel-rooflines $mode $extra > gpurun_out/r02_bench_$tag.log 2>&1
  echo "bench $tag rc=$?"
  python - "$tag" <<'PY'
import json, sys
tag = sys.argv[1]
for l in open(f"gpurun_out/r02_bench_{tag}.log"):
    if l.startswith("{"):
        d = json.loads(l)
        k1 = [o for o in d["roofline_all"] if o["op"].startswith("K1")]
        print("%s: value %.1f clips/s %.2f ms/step  e2e %.1f clips/s  queue=%s loss=%.4f  K1 in-step %.1f us/step in %.0f launches" % (
            tag, d["value"], d["ms_per_step"], d["e2e"]["value"], d["config"].get("queue"), d["loss"],
            1e3 * k1[0]["step_share_ms"], k1[0]["launches_per_step"]))
PY
  grep -v '^{' gpurun_out/r02_bench_$tag.log | grep -i "error\|Traceback" -A5 | head -12
done
