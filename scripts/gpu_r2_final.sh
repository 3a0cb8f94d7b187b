#!/bin/bash
# round 2, final validation on one B200: the whole GPU test suite, smoke, the default bench invocation
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r02_pytest_gpu_final.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/r02_pytest_gpu_final.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke.log
timeout 900 python bench.py > gpurun_out/r02_bench_final.log 2>&1; echo "bench rc=$?"
python - <<'PY'
import json
for l in open("gpurun_out/r02_bench_final.log"):
    if l.startswith("{"):
        d = json.loads(l)
        print("value %.1f clips/s %.2f ms/step e2e %.1f (%.2f ms) steps=%d warmup=%d launches=%d clocks=%s" % (
            d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["steps"], d["warmup"], d["gpu_launches"], d["clocks"]))
        r = d["roofline"]
        print("roofline:", r["kernel"], "frac %.3f traffic %s us %.1f as3 %.3f" % (r["frac"], r["traffic"], r["us_per_instance"], r["as_three_reference_ops"]["frac"]))
        for o in d["roofline_all"]:
            print("  %-40s %7.1f us/inst  %.3f" % (o["op"], o["us_per_instance"], o.get("frac_hbm", 0)))
PY
grep -v '^{' gpurun_out/r02_bench_final.log | grep -i "error\|Traceback" -A5 | head
