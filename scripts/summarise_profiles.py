"""Turn gpurun_out/ ncu artefacts into the tracked summaries under profiles/.

    python scripts/summarise_profiles.py r01

Reads gpurun_out/launches.csv (ncu --metrics gpu__time_duration.sum launch list of a bench.py run)
and gpurun_out/prof_kernels.ncu-rep (ncu --set full capture of our kernels); writes
profiles/<tag>_launch_shares.txt, profiles/<tag>_kernels_ncu.csv and profiles/traffic.json
(dram bytes per launch of each kernel, consumed by bench.py's roofline.traffic).
"""
import collections
import csv
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "profiles")
GP = os.path.join(ROOT, "gpurun_out")
OURS = ("infonce", "ema_multi", "fra_", "hw_mean", "enqueue_kernel", "lmcl_kernel", "queue_transpose", "gather_rows",
        "clip_sgd_multi", "grad_sqnorm_multi", "grad_norm_finish", "color_pipeline", "clip_gray_sum", "flow_visualize",
        "upsample_trilinear", "linear_axis_bwd", "fetch_host")
ENTRY = {"infonce_fused_kernel": "mscl_infonce_fused", "infonce_bwd_slabs_kernel": "mscl_infonce_bwd_slabs", "infonce_bwd_slabs_multi_kernel": "mscl_infonce_bwd_slabs",
         "infonce_tc_kernel": "mscl_infonce_partial", "ema_multi_kernel": "mscl_ema_multi", "fra_maxrad_kernel": "mscl_fra_maxrad",
         "fra_apply_kernel": "mscl_fra_apply", "fra_fused_kernel": "mscl_fra_fused", "hw_mean_fwd_kernel": "mscl_hw_mean_fwd",
         "hw_mean_bwd_kernel": "mscl_hw_mean_bwd", "hw_mean_fwd_small_kernel": "mscl_hw_mean_fwd",
         "hw_mean_bwd_small_kernel": "mscl_hw_mean_bwd", "enqueue_kernel": "mscl_enqueue", "lmcl_kernel": "mscl_lmcl", "clip_sgd_multi_kernel": "mscl_clip_sgd_multi",
         "grad_sqnorm_multi_kernel": "mscl_grad_sqnorm_multi", "color_pipeline_fast_kernel": "mscl_color_pipeline",
         "color_pipeline_kernel": "mscl_color_pipeline", "flow_visualize_kernel": "mscl_flow_visualize",
         "upsample_trilinear_fwd_kernel": "mscl_upsample_trilinear_fwd", "upsample_trilinear_bwd_kernel": "mscl_upsample_trilinear_bwd",
         "upsample_trilinear_ndhwc_fwd_kernel": "mscl_upsample_trilinear_ndhwc_fwd",
         "upsample_trilinear_ndhwc_bwd_kernel": "mscl_upsample_trilinear_ndhwc_bwd",
         "hw_mean_ndhwc_fwd_kernel": "mscl_hw_mean_ndhwc_fwd", "hw_mean_ndhwc_bwd_kernel": "mscl_hw_mean_ndhwc_bwd"}


def launch_shares(tag):
    path = os.path.join(GP, "launches.csv")
    if not os.path.exists(path):
        return
    lines = [l for l in open(path) if not l.startswith("==")]
    r = csv.reader(lines)
    hdr = next(r)
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    tot, cnt = collections.Counter(), collections.Counter()
    for row in r:
        if len(row) <= vi:
            continue
        v = float(row[vi].replace(",", ""))
        v = {"ns": v / 1e3, "us": v, "usecond": v, "ms": v * 1e3}.get(row[ui], v)
        name = re.sub(r"\(.*", "", row[ki])
        name = re.sub(r"^void ", "", name)[:100]
        tot[name] += v
        cnt[name] += 1
    T = sum(tot.values())
    mine = sum(v for k, v in tot.items() if any(o in k for o in OURS))
    with open(os.path.join(OUT, f"{tag}_launch_shares.txt"), "w") as f:
        f.write("# ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off  python bench.py --steps 2 --warmup 3 --no-graphs --profile-range  (scripts/gpu_profile.sh)\n")
        f.write("# per-launch times are cold-cache and serialised: read the SHARES, not the absolutes\n")
        f.write(f"# {sum(cnt.values())} launches, {T/1e3:.2f} ms summed; this repo's kernels: {mine/1e3:.3f} ms = {100*mine/T:.2f} %\n")
        f.write("#   time_us  share%  launches  kernel\n")
        for k, v in tot.most_common(60):
            star = "*" if any(o in k for o in OURS) else " "
            f.write(f"{v:11.1f} {100*v/T:6.2f} {cnt[k]:8d} {star} {k}\n")
        f.write("# --- this repo's kernels (marked * above), all of them ---\n")
        for k, v in tot.most_common():
            if any(o in k for o in OURS):
                f.write(f"{v:11.1f} {100*v/T:6.2f} {cnt[k]:8d} * {k}  ({v/cnt[k]:.1f} us/launch)\n")


def kernel_metrics(tag):
    rep = os.path.join(GP, "prof_kernels.ncu-rep")
    raw_csv = os.path.join(GP, "prof_kernels_raw.csv")       # exported on the GPU box (the .ncu-rep may be too big to travel)
    if os.path.exists(rep):
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    elif os.path.exists(raw_csv):
        raw = open(raw_csv).read()
    else:
        return
    r = list(csv.reader(raw.splitlines()))
    hdr, units, rows = r[0], r[1], r[2:]
    want = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
            "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
            "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
            "launch__shared_mem_per_block_dynamic", "lts__t_sector_hit_rate.pct"]
    idx = [hdr.index(w) for w in want if w in hdr]
    traffic = {}
    with open(os.path.join(OUT, f"{tag}_kernels_ncu.csv"), "w", newline="") as f:
        w = csv.writer(f)
        w.writerow([f"{hdr[i]} [{units[i]}]" if units[i] else hdr[i] for i in idx])
        for row in rows:
            name = re.sub(r"\(.*", "", row[idx[0]]).replace("void ", "").replace("mscl::", "").replace("tc::", "")
            w.writerow([name] + [row[i] for i in idx[1:]])
            for k, entry in ENTRY.items():
                if k in name:
                    def mb(col):
                        i = hdr.index(col)
                        v = float(row[i].replace(",", ""))
                        return v * {"Mbyte": 1e6, "Kbyte": 1e3, "Gbyte": 1e9, "byte": 1}.get(units[i], 1)
                    tr = mb("dram__bytes_read.sum") + mb("dram__bytes_write.sum")
                    grid = row[hdr.index("Grid Size")]
                    traffic.setdefault(entry, {}).setdefault(grid, []).append(tr)
    # one figure per entry point: the launch class with the most traffic (the big launch), averaged
    out = {}
    for entry, by_grid in traffic.items():
        best = max(by_grid.values(), key=lambda v: sum(v) / len(v))
        out[entry] = sum(best) / len(best)
    # per OP (what bench.py's roofline reports): one instance of the K1 op = its forward launch + the backward kernel
    if "mscl_infonce_fused" in out:
        out["K1 InfoNCE (op)"] = out["mscl_infonce_fused"] + out.get("mscl_infonce_bwd_slabs", 0.0)
    for op, entry in (("K4 momentum EMA (op)", "mscl_ema_multi"), ("K3 FRA (op)", "mscl_fra_fused"), ("K5 enqueue (op)", "mscl_enqueue"),
                      ("K8 flow visualiser (op, adjacent)", "mscl_flow_visualize"), ("K9 colour pipeline (op, adjacent)", "mscl_color_pipeline")):
        if entry in out:
            out[op] = out[entry]
    with open(os.path.join(OUT, "traffic.json"), "w") as f:
        json.dump(out, f, indent=1)
    print(out)


if __name__ == "__main__":
    tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
    os.makedirs(OUT, exist_ok=True)
    launch_shares(tag)
    kernel_metrics(tag)
