"""Per-kernel roofline probe: every kernel of the MSCL hot path timed ALONE at BASELINE.json's shapes.

bench.py times whole steps; its in-step per-launch events include ~3 us of launch gap, which
swamps kernels that run for 5-20 us.  Here each kernel is launched back to back over a ring of
distinct buffer sets whose total footprint exceeds twice the L2 (so every launch streams from
HBM), with one pair of CUDA events around the whole train on the launching stream:

    us per launch = elapsed / launches;   achieved = algorithmic bytes / us;   frac = achieved / measured HBM peak

Algorithmic bytes per launch are the ones `functional.py` states.

    python -m mscl_b200.kernel_bench [--configs cfg2,cfg3,cfg4,cfg5] [--out gpurun_out/kernel_rooflines.json]
"""
import argparse
import json
import math
import os

import numpy as np
import torch
import torch.nn.functional as F

from . import _cabi, functional as fx

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
L2_BYTES = 126 * 1024 * 1024


def hbm_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "MEASURED_PEAKS.json"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def _st():
    return torch.cuda.current_stream().cuda_stream      # evaluated per call: the capturing stream during graph capture


def n_rot(bytes_per_set):
    return max(2, min(64, -(-2 * L2_BYTES // max(int(bytes_per_set), 1))))


def time_train(fn, iters, warm=3):
    """us per launch of `fn(i)`, i = 0..iters-1, replayed from ONE CUDA graph so the host (Python, ctypes,
    tensor-map encoding) is out of the measurement: what remains is device time, launch gaps included."""
    for i in range(warm):
        fn(i)
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        with torch.cuda.graph(graph, stream=side):
            for i in range(iters):
                fn(i)
    torch.cuda.current_stream().wait_stream(side)
    graph.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    graph.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / iters


def row(cfg, kernel, shape, us, algo_bytes, algo_flops, pk, bound="hbm", note=None):
    r = dict(config=cfg, kernel=kernel, shape=shape, us=us, algo_bytes=int(algo_bytes), algo_flops=int(algo_flops),
             bound=bound)
    if algo_bytes:
        r["gbs"] = algo_bytes / us / 1e3
        r["frac_hbm"] = r["gbs"] / pk
    if algo_flops:
        r["tflops"] = algo_flops / us / 1e6
    if note:
        r["note"] = note
    return r


# ------------------------------------------------------------------------------------------ K1
def bench_k1(cfg, M, K, pk, dev, iters=30):
    rot = n_rot(K * 512)
    g = torch.Generator().manual_seed(K + M)
    queues = []
    for s in range(rot):
        nq = fx.NegativeQueue(K, 128, dev)
        nq.load(F.normalize(torch.randn(128, K, generator=g), dim=0), torch.randint(0, 2000, (K,), generator=g), 0)
        queues.append(nq)
    q = F.normalize(torch.randn(M, 128, generator=g), dim=1).to(dev)
    k = F.normalize(torch.randn(M, 128, generator=g), dim=1).to(dev)
    qpack = torch.empty(M, fx.PACK_LD, device=dev)
    dscales = [torch.empty((K + 127) // 128 * 128, device=dev) for _ in queues]
    n_part = _cabi.query("mscl_infonce_num_partials", M, K, fx.sm_count(dev))
    part = torch.empty(n_part, M, fx.PACK_LD, device=dev)
    row_loss, dq, gout = torch.empty(2 * M, device=dev), torch.empty(M, 128, device=dev), torch.empty(1, 4, device=dev)

    def prep(i):
        nq = queues[i % rot]
        _cabi.call("mscl_infonce_prep", q.data_ptr(), k.data_ptr(), M, nq.birth.data_ptr(), nq.qstate.data_ptr(), K, 1 / 0.07,
                   1.0, qpack.data_ptr(), dscales[i % rot].data_ptr(), None, 1, None, 0, 0, _st())

    def partial(i):
        _cabi.call("mscl_infonce_partial", qpack.data_ptr(), M, queues[i % rot].queue_tf32.data_ptr(),
                   dscales[i % rot].data_ptr(), K, 0, part.data_ptr(), n_part, 1, _st())

    def finalize(i):
        _cabi.call("mscl_infonce_finalize", qpack.data_ptr(), k.data_ptr(), part.data_ptr(), n_part, M, M, 1 / 0.07, 1,
                   row_loss.data_ptr(), dq.data_ptr(), gout.data_ptr(), _st())

    def whole(i):
        prep(i), partial(i), finalize(i)

    for i in range(rot):
        prep(i)
    ab = fx.infonce_algo_bytes(M, K)
    fl = 4 * M * K * 128
    shape = f"M={M} K={K}"
    out = [row(cfg, "infonce_tc_kernel<grad> (K1 pass, slab form)", shape, time_train(partial, iters), ab, fl, pk,
               note=f"{n_part} CTAs along the keys; queue ring of {rot}"),
           row(cfg, "K1 op, slab form = prep + pass + finalize (3 launches)", shape, time_train(whole, iters), ab, fl, pk)]
    # the single-launch form (csrc/infonce_fused.cu): what functional.infonce runs
    n_fp = _cabi.query("mscl_infonce_fused_parts", M, K, fx.sm_count(dev))
    ws = torch.zeros(16 * M * 4 + 4, device=dev)
    fpart = torch.empty(n_fp, M, fx.PACK_LD, device=dev)
    rowaux = torch.empty(M, 4, device=dev)
    gone = torch.ones(1, device=dev)

    def fused(i, grad=1):
        nq = queues[i % rot]
        _cabi.call("mscl_infonce_fused", q.data_ptr(), k.data_ptr(), M, nq.queue_tf32.data_ptr(), nq.birth.data_ptr(),
                   nq.qstate.data_ptr(), K, 1 / 0.07, 1.0, None, 1, ws.data_ptr(), fpart.data_ptr(), n_fp, M, grad, 1,
                   row_loss.data_ptr(), rowaux.data_ptr(), gout.data_ptr(), _st())

    def bwd(i):
        _cabi.call("mscl_infonce_bwd_slabs", fpart.data_ptr(), n_fp, M, k.data_ptr(), rowaux.data_ptr(), gone.data_ptr(), M,
                   dq.data_ptr(), _st())

    def fused_fb(i):
        fused(i), bwd(i)

    out += [row(cfg, "K1 op = infonce_fused_kernel<grad> (1 launch: prep + pass + loss / top-k)", shape, time_train(fused, iters),
                ab, fl, pk, note=f"{n_fp} CTAs along the keys; the O slabs are summed by the backward kernel"),
            row(cfg, "K1 op + backward = infonce_fused_kernel + infonce_bwd_slabs (2 launches)", shape, time_train(fused_fb, iters),
                ab, fl, pk),
            row(cfg, "K1 op, forward only (no grad)", shape, time_train(lambda i: fused(i, 0), iters), ab, 2 * M * K * 128, pk)]
    del queues
    return out


def bench_k1_pair(cfg, Ms, K, pk, dev, iters=30):
    """What MSCLWithAug.objective launches for its two independent pre-enqueue passes: len(Ms) jobs over different queues
    in ONE mscl_infonce_fused_multi launch (disjoint SMs, shared ramp / drain / finalize)."""
    import ctypes
    n = len(Ms)
    rot = n_rot(n * K * 512)
    g = torch.Generator().manual_seed(K + sum(Ms))
    rings = []
    for s in range(rot):
        qs = []
        for _ in range(n):
            nq = fx.NegativeQueue(K, 128, dev)
            nq.load(F.normalize(torch.randn(128, K, generator=g), dim=0), torch.randint(0, 2000, (K,), generator=g), 0)
            qs.append(nq)
        rings.append(qs)
    q = [F.normalize(torch.randn(M, 128, generator=g), dim=1).to(dev) for M in Ms]
    k = [F.normalize(torch.randn(M, 128, generator=g), dim=1).to(dev) for M in Ms]
    arr_i32 = lambda v: (ctypes.c_int32 * n)(*v)
    arr_i64 = lambda v: (ctypes.c_int64 * n)(*v)
    arr_f32 = lambda v: (ctypes.c_float * n)(*v)
    arr_ptr = lambda ts: (ctypes.c_void_p * n)(*[(t.data_ptr() if t is not None else None) for t in ts])
    n_part = _cabi.query("mscl_infonce_fused_parts_multi", n, arr_i32(Ms), arr_i64([K] * n), fx.sm_count(dev))
    ws = [torch.zeros(16 * M * 4 + 4, device=dev) for M in Ms]
    part = [torch.empty(n_part, M, fx.PACK_LD, device=dev) for M in Ms]
    rowaux = [torch.empty(M, 4, device=dev) for M in Ms]
    row_loss = [torch.empty(2 * M, device=dev) for M in Ms]
    gout = [torch.empty(1, 4, device=dev) for _ in Ms]
    dq = [torch.empty(M, 128, device=dev) for M in Ms]
    gone = torch.ones(1, device=dev)
    # every table is built once: the launches of the train only differ in the queue ring slot
    tabs = [dict(queue=arr_ptr([nq.queue_tf32 for nq in qs]), birth=arr_ptr([nq.birth for nq in qs]),
                 qstate=arr_ptr([nq.qstate for nq in qs])) for qs in rings]
    fixed = dict(q=arr_ptr(q), k=arr_ptr(k), M=arr_i32(Ms), K=arr_i64([K] * n), invT=arr_f32([1 / 0.07] * n), bound=arr_f32([1.0] * n),
                 dup=arr_ptr([None] * n), age=arr_i32([1] * n), ws=arr_ptr(ws), part=arr_ptr(part), rpg=arr_i32(Ms),
                 flags=arr_i32([1] * n), rl=arr_ptr(row_loss), ra=arr_ptr(rowaux), go=arr_ptr(gout))

    def fwd(i):
        t = tabs[i % rot]
        _cabi.call("mscl_infonce_fused_multi", n, fixed["q"], fixed["k"], fixed["M"], t["queue"], t["birth"], t["qstate"], fixed["K"],
                   fixed["invT"], fixed["bound"], fixed["dup"], fixed["age"], fixed["ws"], fixed["part"], n_part, fixed["rpg"], 1,
                   fixed["flags"], fixed["rl"], fixed["ra"], fixed["go"], _st())

    gones = [gone] * n

    def fwd_bwd(i):
        fwd(i)
        _cabi.call("mscl_infonce_bwd_slabs_multi", n, fixed["part"], n_part, fixed["M"], fixed["k"], fixed["ra"], arr_ptr(gones),
                   fixed["rpg"], arr_ptr(dq), _st())

    ab = sum(fx.infonce_algo_bytes(M, K) for M in Ms)
    fl = sum(4 * M * K * 128 for M in Ms)
    shape = "M=" + "+".join(str(M) for M in Ms) + f" K={K} x{n} queues"
    out = [row(cfg, f"K1 x{n} ops in one launch = infonce_fused_kernel<grad>, {n} jobs (what the step runs)", shape,
               time_train(fwd, iters), ab, fl, pk, note=f"{n_part} CTAs per job"),
           row(cfg, f"K1 x{n} ops in one launch + their backward kernel", shape, time_train(fwd_bwd, iters), ab, fl, pk)]
    del rings
    return out


def bench_k1_step(cfg, n, K, pk, dev, iters=30):
    """The ONE launch MSCLWithAug.objective makes for all seven InfoNCE terms of a step (mscl_infonce_fused_multi_x):
    job 0 = 3n rows over W_rgb; job 1 = 4n rows over W_flow with an epoch split -- n rows read it as it was before the
    base-flow enqueue (ages - 1, the n slots it wrote still holding the keys it overwrote), 3n rows as it is (n of them
    finding their own positive among the new keys).  The reference's three distinct negative matrices cost two queue
    reads."""
    import ctypes
    Ms = [3 * n, 4 * n]
    rot = n_rot(2 * K * 512)
    g = torch.Generator().manual_seed(K + n)
    rings = []
    for s in range(rot):
        qs = []
        for _ in range(2):
            nq = fx.NegativeQueue(K, 128, dev)
            nq.load(F.normalize(torch.randn(128, K, generator=g), dim=0), torch.randint(0, 2000, (K,), generator=g), 5 * n)
            qs.append(nq)
        rings.append(qs)
    q = [F.normalize(torch.randn(M, 128, generator=g), dim=1).to(dev) for M in Ms]
    k = [F.normalize(torch.randn(M, 128, generator=g), dim=1).to(dev) for M in Ms]
    xkeys = F.normalize(torch.randn(n, 128, generator=g), dim=1).to(dev)       # what the enqueue overwrote ...
    xbirth = torch.randint(0, 100, (n,), generator=g, dtype=torch.int32).to(dev)      # ... and its births
    dup = torch.full((4 * n,), -1, dtype=torch.int32)
    dup[2 * n:3 * n] = torch.arange(n, dtype=torch.int32) + 5 * n
    dup = dup.to(dev)
    arr_i32 = lambda v: (ctypes.c_int32 * 2)(*v)
    arr_i64 = lambda v: (ctypes.c_int64 * 2)(*v)
    arr_f32 = lambda v: (ctypes.c_float * 2)(*v)
    arr_ptr = lambda ts: (ctypes.c_void_p * 2)(*[(t.data_ptr() if t is not None else None) for t in ts])
    n_part = _cabi.query("mscl_infonce_fused_parts_multi", 2, arr_i32(Ms), arr_i64([K] * 2), fx.sm_count(dev))
    ws = [torch.zeros(16 * M * 4 + 4, device=dev) for M in Ms]
    part = [torch.empty(n_part, M, fx.PACK_LD, device=dev) for M in Ms]
    rowaux = [torch.empty(M, 4, device=dev) for M in Ms]
    row_loss = [torch.empty(2 * M, device=dev) for M in Ms]
    gout = [torch.empty(M // n, 4, device=dev) for M in Ms]
    dq = [torch.empty(M, 128, device=dev) for M in Ms]
    gone = [torch.ones(M // n, device=dev) for M in Ms]
    tabs = [dict(queue=arr_ptr([nq.queue_tf32 for nq in qs]), birth=arr_ptr([nq.birth for nq in qs]),
                 qstate=arr_ptr([nq.qstate for nq in qs])) for qs in rings]
    fixed = dict(q=arr_ptr(q), k=arr_ptr(k), M=arr_i32(Ms), K=arr_i64([K] * 2), invT=arr_f32([1 / 0.07] * 2), bound=arr_f32([1.0] * 2),
                 dup=arr_ptr([None, dup]), age=arr_i32([1, 1]), ws=arr_ptr(ws), part=arr_ptr(part), rpg=arr_i32([n, n]),
                 flags=arr_i32([1, 1]), rl=arr_ptr(row_loss), ra=arr_ptr(rowaux), go=arr_ptr(gout))

    def fwd(i):
        t = tabs[i % rot]
        _cabi.call("mscl_infonce_fused_multi_x", 2, fixed["q"], fixed["k"], fixed["M"], t["queue"], t["birth"], t["qstate"],
                   fixed["K"], fixed["invT"], fixed["bound"], fixed["dup"], fixed["age"], fixed["ws"], fixed["part"], n_part,
                   fixed["rpg"], 1, fixed["flags"], fixed["rl"], fixed["ra"], fixed["go"], 1, xkeys.data_ptr(), xbirth.data_ptr(), 5 * n, n, n,
                   _st())

    def fwd_bwd(i):
        fwd(i)
        _cabi.call("mscl_infonce_bwd_slabs_multi", 2, fixed["part"], n_part, fixed["M"], fixed["k"], fixed["ra"], arr_ptr(gone),
                   fixed["rpg"], arr_ptr(dq), _st())

    ab = sum(fx.infonce_algo_bytes(M, K) for M in Ms) + n * 512
    fl = sum(4 * M * K * 128 for M in Ms)
    ab3 = fx.infonce_algo_bytes(3 * n, K) * 2 + fx.infonce_algo_bytes(n, K)       # the same terms as three separate ops
    shape = f"M={Ms[0]}+({n}|{3 * n}) K={K} x2 queues"
    us_f, us_fb = time_train(fwd, iters), time_train(fwd_bwd, iters)
    out = [row(cfg, "K1 step launch = infonce_fused_kernel<grad>, 2 jobs, epoch split on W_flow (all 7 terms of a step)", shape,
               us_f, ab, fl, pk, note=f"{n_part} CTAs per job; bytes = the two queues once each; the same terms as three "
               f"separate ops are {ab3} algorithmic bytes = {ab3 / us_f / 1e3 / pk:.3f} of the HBM roofline, {us_f / 3:.1f} us per op"),
           row(cfg, "K1 step launch + its backward kernel", shape, us_fb, ab, fl, pk)]
    del rings
    return out


# ------------------------------------------------------------------------------------------ K2
def bench_k2(cfg, N, t, pk, dev, iters=30):
    out = []
    for name, shp in (("rgb", (N, 128, t, 28, 28)), ("flow", (N, 128, 2 * t, 7, 7))):
        nbytes = 4 * int(np.prod(shp))
        rot = n_rot(nbytes)
        xs = [torch.randn(shp, device=dev) for _ in range(rot)]
        HW = shp[-1] * shp[-2]
        R = xs[0].numel() // HW
        o = torch.empty(shp[:3], device=dev)
        us = time_train(lambda i: _cabi.call("mscl_hw_mean_fwd", xs[i % rot].data_ptr(), o.data_ptr(), R, HW, _st()), iters)
        out.append(row(cfg, "hw_mean_fwd", f"{name} {shp}", us, 4 * R * (HW + 1), 0, pk))
        us = time_train(lambda i: _cabi.call("mscl_hw_mean_bwd", o.data_ptr(), xs[i % rot].data_ptr(), R, HW, _st()), iters)
        out.append(row(cfg, "hw_mean_bwd", f"{name} {shp}", us, 4 * R * (HW + 1), 0, pk))
        # channels-last kernels on the same buffers viewed as [N, T, HW, C] (what the channels-last step runs)
        n_, c_, t_ = shp[:3]
        us = time_train(lambda i: _cabi.call("mscl_hw_mean_ndhwc_fwd", xs[i % rot].data_ptr(), o.data_ptr(), n_, c_, t_, HW, _st()), iters)
        out.append(row(cfg, "hw_mean_ndhwc_fwd", f"{name} {shp}", us, 4 * R * (HW + 1), 0, pk))
        us = time_train(lambda i: _cabi.call("mscl_hw_mean_ndhwc_bwd", o.data_ptr(), xs[i % rot].data_ptr(), n_, c_, t_, HW, _st()), iters)
        out.append(row(cfg, "hw_mean_ndhwc_bwd", f"{name} {shp}", us, 4 * R * (HW + 1), 0, pk))
        del xs
    xq = torch.randn(N, 128, t, device=dev)
    xf = torch.randn(N, 128, 2 * t, device=dev)
    us = time_train(lambda i: fx._LMCL.apply(xq, xf, 1 / 0.07), iters)
    out.append(row(cfg, "lmcl_kernel (fwd+bwd)", f"N={N} t={t}", us, 8 * N * 128 * 3 * t, 6 * N * t * 2 * t * 128, pk,
                   bound="latency", note="one CTA per clip, KFLOPs: launch-latency bound"))
    return out


# ------------------------------------------------------------------------------------------ K3
def bench_k3(cfg, N, T, pk, dev, iters=30):
    HW = 112 * 112
    in_bytes = N * 2 * T * HW * 4
    rot = n_rot(3 * in_bytes)
    flows = [torch.randn(N, 2, T, 112, 112, device=dev) for _ in range(rot)]
    outs = [torch.empty(N, 2, 2 * T, 112, 112, device=dev) for _ in range(rot)]
    cid = torch.from_numpy(np.random.RandomState(0).randint(0, 8, size=N).astype(np.int32)).to(dev)
    tab = fx.fra_table(device=dev)
    maxrad = torch.empty(N, T, 2, device=dev)
    us1 = time_train(lambda i: _cabi.call("mscl_fra_maxrad", flows[i % rot].data_ptr(), cid.data_ptr(), tab.data_ptr(),
                                          maxrad.data_ptr(), N, T, HW, 0, _st()), iters)
    us2 = time_train(lambda i: _cabi.call("mscl_fra_apply", flows[i % rot].data_ptr(), cid.data_ptr(), tab.data_ptr(),
                                          maxrad.data_ptr(), outs[i % rot].data_ptr(), N, T, HW, 0, _st()), iters)
    us3 = time_train(lambda i: _cabi.call("mscl_fra_fused", flows[i % rot].data_ptr(), cid.data_ptr(), tab.data_ptr(),
                                          outs[i % rot].data_ptr(), N, T, HW, 0, _st()), iters)
    shape = f"({N},2,{T},112,112)"
    return [row(cfg, "fra_fused_kernel (one pass)", shape, us3, 24 * N * T * HW, 0, pk),
            row(cfg, "fra_maxrad", shape, us1, 8 * N * T * HW, 0, pk, note="includes the 2-float-per-frame memset"),
            row(cfg, "fra_apply", shape, us2, 24 * N * T * HW, 0, pk, note="two-pass form, frames too large for a cluster")]


# ------------------------------------------------------------------------------------------ K4
def r3d18_key_sizes():
    """Parameter sizes of the r18 RGB key side (R3D-18 + TPN neck + MLP): 36.7 M elements."""
    sizes = [3 * 64 * 3 * 49, 64, 64]
    cin = 64
    for planes, blocks in ((64, 2), (128, 2), (256, 2), (512, 2)):
        for b in range(blocks):
            sizes += [cin * planes * 27, planes, planes, planes * planes * 27, planes, planes]
            if b == 0 and cin != planes:
                sizes += [cin * planes, planes, planes]
            cin = planes
    sizes += [128 * 128, 128, 256 * 128, 128, 512 * 128, 128] + [128 * 128 * 9, 128] * 3 + [128 * 128 * 27, 128] * 3
    sizes += [512 * 512, 512, 512 * 128, 128]
    return sizes


def slowonly_r50_key_sizes():
    """SlowOnly-R50 key side, cfg 5: 159 backbone tensors (min 64, max 3,145,728 elements) + neck + MLP = 38.4 M."""
    sizes = [3 * 64 * 49, 64, 64]
    cin = 64
    for planes, blocks, kt in ((64, 3, 1), (128, 4, 1), (256, 6, 3), (512, 3, 3)):
        for b in range(blocks):
            sizes += [cin * planes * kt, planes, planes, planes * planes * 9, planes, planes, planes * planes * 4,
                      planes * 4, planes * 4]
            if b == 0:
                sizes += [cin * planes * 4, planes * 4, planes * 4]
            cin = planes * 4
    sizes += [512 * 128, 128, 1024 * 128, 128, 2048 * 128, 128] + [128 * 128 * 9, 128] * 3 + [128 * 128 * 27, 128] * 3
    sizes += [2048 * 2048, 2048, 2048 * 128, 128]
    return sizes


def bench_k4(cfg, name, sizes, pk, dev, iters=20):
    total = sum(sizes)
    rot = n_rot(8 * total)
    tables = []
    for _ in range(rot):
        flat_k, flat_q = torch.randn(total, device=dev), torch.randn(total, device=dev)
        ks, qs, o = [], [], 0
        for n in sizes:       # views into one allocation per side: same access pattern as separate 256B-aligned tensors
            ks.append(flat_k[o:o + n]), qs.append(flat_q[o:o + n])
            o += n
        tables.append(fx.EmaTable(ks, qs))
    us = time_train(lambda i: tables[i % rot].update(0.997), iters)
    return [row(cfg, "ema_multi_kernel", f"{name}: {total} elements / {len(sizes)} tensors", us, 12 * total, 3 * total, pk)]


# ------------------------------------------------------------------------------------------ K5 / K6
def bench_k5(cfg, B_all, K, pk, dev, iters=30):
    nq = fx.NegativeQueue(K, 128, dev)
    nq.load(F.normalize(torch.randn(128, K), dim=0), torch.zeros(K, dtype=torch.long), 0)
    keys = F.normalize(torch.randn(B_all, 128, device=dev), dim=1)
    us = time_train(lambda i: nq.enqueue(keys), iters)
    return [row(cfg, "enqueue_kernel", f"B_all={B_all} K={K}", us, 2 * B_all * 128 * 4, 0, pk, bound="latency",
                note="64-256 KB per call: launch-latency bound")]


def bench_k6(cfg, shape, pk, dev, iters=20):
    nbytes = 4 * int(np.prod(shape))
    rot = n_rot(2 * nbytes)
    xs = [torch.randn(shape, device=dev) for _ in range(rot)]
    idx = torch.randperm(shape[0]).to(dev)
    us = time_train(lambda i: fx.gather_rows(xs[i % rot], idx), iters)
    return [row(cfg, "gather_rows_kernel (shuffle-BN)", str(tuple(shape)), us, 2 * nbytes, 0, pk)]


# ------------------------------------------------------------------------------------------ K7 / K8 / K9
def bench_k789(cfg, N, pk, dev, iters=20):
    out = []
    # K7: the two SEPC up-sampling steps of the r18 TPN (necks/sepc.py:126-130)
    for shp, size in (((N, 128, 2, 14, 14), (4, 28, 28)), ((N, 128, 1, 7, 7), (2, 14, 14))):
        nout = N * 128 * size[0] * size[1] * size[2]
        nin = int(np.prod(shp))
        rot = n_rot(4 * (nin + nout))
        xs = [torch.randn(shp, device=dev) for _ in range(rot)]
        ys = [torch.empty((N, 128) + size, device=dev) for _ in range(rot)]
        us = time_train(lambda i: _cabi.call("mscl_upsample_trilinear_fwd", xs[i % rot].data_ptr(), ys[i % rot].data_ptr(),
                                             N * 128, *shp[2:], *size, _st()), iters)
        out.append(row(cfg, "upsample_trilinear_fwd", f"{shp} -> {size}", us, 4 * (nin + nout), 0, pk))
        us = time_train(lambda i: _cabi.call("mscl_upsample_trilinear_bwd", ys[i % rot].data_ptr(), xs[i % rot].data_ptr(),
                                             N * 128, *shp[2:], *size, _st()), iters)
        out.append(row(cfg, "upsample_trilinear_bwd", f"{size} -> {shp}", us, 4 * (nin + nout), 0, pk))
        # the channels-last kernels (what the step runs: the encoders are in channels_last_3d); same buffers, NDHWC view
        us = time_train(lambda i: _cabi.call("mscl_upsample_trilinear_ndhwc_fwd", xs[i % rot].data_ptr(), ys[i % rot].data_ptr(),
                                             N, 128, *shp[2:], *size, _st()), iters)
        out.append(row(cfg, "upsample_trilinear_ndhwc_fwd", f"{shp} -> {size}", us, 4 * (nin + nout), 0, pk))
        us = time_train(lambda i: _cabi.call("mscl_upsample_trilinear_ndhwc_bwd", ys[i % rot].data_ptr(), xs[i % rot].data_ptr(),
                                             N, 128, *shp[2:], *size, _st()), iters)
        out.append(row(cfg, "upsample_trilinear_ndhwc_bwd", f"{size} -> {shp}", us, 4 * (nin + nout), 0, pk))
        # what the autograd op runs: the separable backward, one 1-D gather pass per scaled axis (W, H, T)
        (ti, hi, wi), (to, ho, wo) = shp[2:], size
        t1 = torch.empty(N * to * ho * wi * 128, device=dev)
        t2 = torch.empty(N * to * hi * wi * 128, device=dev)

        def separable(i):
            cur = ys[i % rot].data_ptr()
            if wi != wo:
                _cabi.call("mscl_linear_axis_bwd", cur, t1.data_ptr(), N * to * ho, wi, wo, 32, _st())
                cur = t1.data_ptr()
            if hi != ho:
                _cabi.call("mscl_linear_axis_bwd", cur, t2.data_ptr(), N * to, hi, ho, wi * 32, _st())
                cur = t2.data_ptr()
            if ti != to:
                _cabi.call("mscl_linear_axis_bwd", cur, xs[i % rot].data_ptr(), N, ti, to, hi * wi * 32, _st())
        us = time_train(separable, iters)
        out.append(row(cfg, "linear_axis_bwd x3 (separable K7 bwd)", f"{size} -> {shp}", us, 4 * (nin + nout), 0, pk))
        del xs, ys, t1, t2
    # K8: (N,2,16,112,112) flow -> colour image, K9: (N,3,8,112,112) RGB clips
    T2, HW = 16, 112 * 112
    rot = n_rot(20 * N * T2 * HW)
    flows = [torch.randn(N, 2, T2, 112, 112, device=dev) for _ in range(rot)]
    imgs = [torch.empty(N, 3, T2, 112, 112, device=dev) for _ in range(rot)]
    flip = (torch.arange(N, device=dev) % 2).to(torch.uint8)
    us = time_train(lambda i: _cabi.call("mscl_flow_visualize", flows[i % rot].data_ptr(), flip.data_ptr(), None,
                                         imgs[i % rot].data_ptr(), N, T2, 112, 112, _st()), iters)
    out.append(row(cfg, "flow_visualize_kernel", f"({N},2,{T2},112,112)", us, 20 * N * T2 * HW, 0, pk,
                   note="atan2 + float64 interpolation per pixel: ALU-heavy for an element-wise pass"))
    del flows, imgs
    from .common.ssl_aug import SyncMoCoAugmentV5
    aug = SyncMoCoAugmentV5(crop_size=112, sync_level=("batch", "batch"), t=(8, 8), flow_suffix="flow_imgs")
    T = 8
    rot = n_rot(24 * N * T * HW)
    xs = [torch.rand(N, 3, T, 112, 112, device=dev) for _ in range(rot)]
    ys = [torch.empty(N, 3, T, 112, 112, device=dev) for _ in range(rot)]
    torch.manual_seed(0)
    norm = torch.cat([aug.mean.view(-1), aug.std.view(-1)]).to(dev)
    # per-frame jitter factors (the config's 'batch' sync level: what the step runs; the frame's mean luminance is formed in
    # the kernel by the cluster of its bands) and per-clip ones ('params' level: clip_gray_sum pre-pass + the pipeline)
    for frames, label in ((T, "per-frame factors: color_pipeline (cluster mean)"), (1, "per-clip factors: clip_gray_sum + color_pipeline")):
        prm = aug._color_params(N, dev, frames)
        chunks = 1 if frames > 1 else fx._GRAY_CHUNKS
        scratch = torch.empty(N * frames, chunks, device=dev)
        cases = (("drawn decisions (p_blur=.5)", None), ("all clips blurred", True), ("no clip blurred", False)) if frames > 1 else \
            (("drawn decisions (p_blur=.5)", None),)
        for name, blur in cases:
            if blur is not None:
                prm["blur"] = torch.full((N * frames,), blur, device=dev)
            params = aug._pack_params(prm, flip.bool().repeat_interleave(frames), False)
            taps = prm["taps"].contiguous()
            us = time_train(lambda i: _cabi.call("mscl_color_pipeline", xs[i % rot].data_ptr(), params.data_ptr(), taps.data_ptr(),
                                                 taps.numel(), norm.data_ptr(), scratch.data_ptr(), chunks,
                                                 ys[i % rot].data_ptr(), N, T, 112, 112, int(frames > 1), _st()), iters)
            out.append(row(cfg, label, f"({N},3,{T},112,112) {name}", us, (24 if frames > 1 else 36) * N * T * HW, 0, pk))
    return out


def run(configs=("cfg2", "cfg3", "cfg4", "cfg5"), device=0, verbose=True):
    dev = torch.device("cuda", device)
    torch.cuda.set_device(dev)
    _cabi.require_device(device)
    pk, src = hbm_peak()
    rows = []

    def add(rs):
        rows.extend(rs)
        if verbose:
            for r in rs:
                frac = f"{100 * r['frac_hbm']:5.1f}% of HBM peak" if "frac_hbm" in r else ""
                print(f"[{r['config']}] {r['kernel']:<78} {r['shape']:<44} {r['us']:8.1f} us  "
                      f"{r.get('gbs', 0):7.0f} GB/s  {frac}", flush=True)
        torch.cuda.empty_cache()

    if "cfg2" in configs:      # the r18 pre-training step, 32 clips/GPU, K = 65536
        for M in (96, 32):
            add(bench_k1("cfg2", M, 65536, pk, dev))
        add(bench_k1_pair("cfg2", (96, 32), 65536, pk, dev))
        add(bench_k1_step("cfg2", 32, 65536, pk, dev))
        add(bench_k2("cfg2", 32, 4, pk, dev))
        add(bench_k3("cfg2", 32, 8, pk, dev))
        add(bench_k4("cfg2", "r18 RGB key side", r3d18_key_sizes(), pk, dev))
        add(bench_k5("cfg2", 32, 65536, pk, dev))
        add(bench_k789("cfg2", 32, pk, dev))
    if "cfg3" in configs:      # queue sweep, N = 64 per GPU
        for K in (16384, 65536, 262144, 1048576):
            add(bench_k1("cfg3", 64, K, pk, dev, iters=20))
    if "cfg4" in configs:      # LMCL + FRA, N = 64, t = 16
        add(bench_k2("cfg4", 64, 16, pk, dev, iters=15))
        add(bench_k3("cfg4", 64, 16, pk, dev, iters=15))
    if "cfg5" in configs:      # large-parameter key encoder
        add(bench_k4("cfg5", "SlowOnly-R50 key side", slowonly_r50_key_sizes(), pk, dev))
        add(bench_k6("cfg5", (32, 3, 8, 224, 224), pk, dev))
        add(bench_k5("cfg5", 256, 65536, pk, dev))
    return dict(hbm_peak_gbs=pk, peak_source=src, timing="back-to-back launches over a buffer ring > 2x L2, one CUDA-event pair",
                rows=rows)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--configs", default="cfg2,cfg3,cfg4,cfg5")
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "kernel_rooflines.json"))
    args = ap.parse_args()
    res = run(tuple(args.configs.split(",")))
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    with open(args.out, "w") as f:
        json.dump(res, f, indent=1)


if __name__ == "__main__":
    main()
