#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
g=$(nvidia-smi -L | wc -l)
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $g --master-addr 127.0.0.1 --master-port $((29500 + g)) \
    bench.py --gpus $g --steps 10 --warmup 4 --no-cpu-baseline --no-kernel-rooflines --timeline-out gpurun_out/r02_timeline_host_g$g.txt > gpurun_out/r02_bench_host_g$g.log 2>&1
echo "rc=$?"; grep -A45 "host side" gpurun_out/r02_timeline_host_g$g.txt | cut -c1-170
timeout 600 python bench.py --gpus 1 --steps 10 --warmup 4 --no-cpu-baseline --no-kernel-rooflines --no-gpu-eager-baseline --timeline-out gpurun_out/r02_timeline_host_g1.txt > gpurun_out/r02_bench_host_g1.log 2>&1
echo "rc=$?"; head -3 gpurun_out/r02_timeline_host_g1.txt | cut -c1-200; grep -A25 "host side" gpurun_out/r02_timeline_host_g1.txt | cut -c1-170
