#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29502 \
    bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline --no-kernel-rooflines --timeline-out gpurun_out/r02_timeline_g2.txt > gpurun_out/r02_bench_g2.log 2>&1
echo "bench g=2 rc=$?"
grep -c "Grad strides do not match" gpurun_out/r02_bench_g2.log
grep -n "Grad strides" -A2 gpurun_out/r02_bench_g2.log | head -8
grep -v '^{' gpurun_out/r02_bench_g2.log | grep -i "error\|Traceback" -A5 | head -20
head -30 gpurun_out/r02_timeline_g2.txt
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29602 \
      scripts/k1_sweep_multi.py --Ks 65536 --Ms 96 > gpurun_out/r02_k1_sweep_g2.log 2>&1; grep "^G=" gpurun_out/r02_k1_sweep_g2.log
