#!/bin/bash
# ncu evidence for the round: launch list of a bench step + full captures of the hot kernels.
# gpurun copies back at most 64 MiB: the .ncu-rep is exported to CSV on the box and dropped if it is large.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
# launch list: skip the warm-up launches, record 2 steps (cold-cache, serialised: shares only).  Eager encoders
# (--no-graphs): ncu serialises kernel nodes of a replayed graph anyway, and capture under ncu is fragile.
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 20000 --csv \
    --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-graphs --no-cpu-baseline --no-kernel-rooflines --profile-range > gpurun_out/bench_under_ncu.log 2>&1
echo "launchlist rc=$?"; tail -3 gpurun_out/bench_under_ncu.log
# full capture of our kernels (regex on kernel names), a few launches each
timeout 1200 ncu --set full --clock-control none \
    -k regex:"infonce_fused_kernel|infonce_bwd_slabs|infonce_tc_kernel|ema_multi_kernel|fra_|hw_mean|enqueue_kernel|lmcl_kernel|clip_sgd_multi|grad_sqnorm_multi|color_pipeline|clip_gray_sum|flow_visualize|upsample_trilinear|linear_axis_bwd|fetch_host" -c 60 --profile-from-start off \
    -o gpurun_out/prof_kernels -f python bench.py --steps 2 --warmup 3 --no-graphs --no-cpu-baseline --no-kernel-rooflines --profile-range > gpurun_out/prof_kernels.log 2>&1
echo "full rc=$?"
ncu -i gpurun_out/prof_kernels.ncu-rep --page raw --csv > gpurun_out/prof_kernels_raw.csv 2>/dev/null
sz=$(stat -c %s gpurun_out/prof_kernels.ncu-rep 2>/dev/null || echo 0)
if [ "$sz" -gt 40000000 ]; then rm -f gpurun_out/prof_kernels.ncu-rep; echo "dropped .ncu-rep ($sz bytes), kept the raw CSV"; fi
du -sh gpurun_out; ls -la gpurun_out
