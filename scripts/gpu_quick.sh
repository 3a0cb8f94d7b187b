#!/bin/bash
# tests + a short bench (no ncu)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 600 python bench.py --steps 10 --warmup 4 --no-cpu-baseline --no-kernel-rooflines > gpurun_out/bench_quick.log 2>&1; echo "bench rc=$?"
grep -v '^{' gpurun_out/bench_quick.log | tail -5
python - <<'PY'
import json
for l in open('gpurun_out/bench_quick.log'):
    if l.startswith('{'):
        d=json.loads(l)
        print("value %.1f clips/s  %.2f ms/step  e2e %.1f clips/s %.2f ms  launches %d loss %.4f"%(d['value'],d['ms_per_step'],d['e2e']['value'],d['e2e']['ms_per_step'],d['gpu_launches'],d['loss']))
        for k in d['kernels']: print("   %-28s n/step=%.1f avg_us=%.1f frac=%s"%(k['kernel'],k['launches_per_step'],k['avg_us'],k.get('frac_hbm')))
PY
