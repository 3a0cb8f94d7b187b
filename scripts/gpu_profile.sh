#!/bin/bash
# ncu evidence for the round: launch list of a bench step + full captures of the hot kernels.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
# launch list: skip the warm-up launches, record ~2 steps (cold-cache, serialised: shares only)
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 20000 --csv \
    --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-kernel-rooflines --profile-range > gpurun_out/bench_under_ncu.log 2>&1
echo "launchlist rc=$?"
# full capture of our kernels (regex on kernel names), a few launches each
timeout 1200 ncu --set full --clock-control none --import-source on \
    -k regex:"infonce_tc_kernel|ema_multi_kernel|fra_|hw_mean|enqueue_kernel|lmcl_kernel|clip_sgd_multi|grad_sqnorm_multi|color_pipeline|flow_visualize|upsample_trilinear" -c 60 --profile-from-start off \
    -o gpurun_out/prof_kernels -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-kernel-rooflines --profile-range > gpurun_out/prof_kernels.log 2>&1
echo "full rc=$?"
ls -la gpurun_out
