#!/usr/bin/env python
"""K1 sweep (BASELINE.json configs[2]): the fused InfoNCE pass alone over K = 16 Ki ... 1 Mi negatives.

Each (M, K) point launches `mscl_infonce_partial` back to back over a ring of distinct queues whose
total size exceeds 2x the L2 (so every launch streams its queue from HBM), timed with one pair of
CUDA events around the whole train of launches -> average device time per launch, algorithmic
GB/s (functional.infonce_algo_bytes) and the fraction of the measured HBM peak.  Also times the
whole op (prep + partial + finalize) the same way.

    python scripts/k1_sweep.py [--out gpurun_out/k1_sweep.json]
"""
import argparse
import json
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mscl_b200 import _cabi, functional as fx  # noqa: E402


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"])
    except Exception:
        return 6650.0


def make_queue(K, dev, seed):
    g = torch.Generator().manual_seed(seed)
    nq = fx.NegativeQueue(K, 128, dev)
    qk = F.normalize(torch.randn(128, K, generator=g), dim=0)
    count = torch.randint(0, 2000, (K,), generator=g)
    nq.load(qk, count, 0)
    return nq


def time_train(fn, n):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n):
        fn(i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / n      # us per launch


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "k1_sweep.json"))
    ap.add_argument("--Ks", default="16384,65536,262144,1048576")
    ap.add_argument("--Ms", default="32,64,96,128,192,384")
    ap.add_argument("--iters", type=int, default=40)
    ap.add_argument("--flags", type=int, default=1, help="mscl_infonce_fused flags (1 = early queue prefetch)")
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    pk = peaks()
    sms = fx.sm_count(dev)
    st = torch.cuda.current_stream().cuda_stream
    rows = []
    for K in [int(x) for x in args.Ks.split(",")]:
        n_rot = max(2, -(-300_000_000 // (K * 512)))
        queues = [make_queue(K, dev, s) for s in range(n_rot)]
        for M in [int(x) for x in args.Ms.split(",")]:
            g = torch.Generator().manual_seed(M)
            q = F.normalize(torch.randn(M, 128, generator=g), dim=1).to(dev)
            k = F.normalize(torch.randn(M, 128, generator=g), dim=1).to(dev)
            qpack = torch.empty(M, fx.PACK_LD, device=dev)
            k_pad = (K + 127) // 128 * 128
            dscales = [torch.empty(k_pad, device=dev) for _ in queues]
            for nq, ds in zip(queues, dscales):
                _cabi.call("mscl_infonce_prep", q.data_ptr(), k.data_ptr(), M, nq.birth.data_ptr(), nq.qstate.data_ptr(),
                           K, 1 / 0.07, 1.0, qpack.data_ptr(), ds.data_ptr(), None, 1, None, 0, 0, st)
            n_part = _cabi.query("mscl_infonce_num_partials", M, K, sms)
            part = torch.empty(n_part, M, fx.PACK_LD, device=dev)
            row_loss = torch.empty(2 * M, device=dev)
            dq = torch.empty(M, 128, device=dev)
            gout = torch.empty(1, 4, device=dev)
            res = dict(M=M, K=K, n_part=n_part, algo_bytes=fx.infonce_algo_bytes(M, K))
            for grad in (1, 0):
                def partial(i, grad=grad):
                    nq = queues[i % n_rot]
                    _cabi.call("mscl_infonce_partial", qpack.data_ptr(), M, nq.queue_tf32.data_ptr(),
                               dscales[i % n_rot].data_ptr(), K, 0, part.data_ptr(), n_part, grad, st)
                for i in range(5):
                    partial(i)
                us = time_train(partial, args.iters)
                tag = "grad" if grad else "nograd"
                res[f"partial_{tag}_us"] = us
                res[f"partial_{tag}_gbs"] = res["algo_bytes"] / us / 1e3
                res[f"partial_{tag}_frac"] = res[f"partial_{tag}_gbs"] / pk
                res[f"partial_{tag}_tflops"] = (4 if grad else 2) * M * K * 128 / us / 1e6

            def whole(i):
                nq = queues[i % n_rot]
                ds = dscales[i % n_rot]
                _cabi.call("mscl_infonce_prep", q.data_ptr(), k.data_ptr(), M, nq.birth.data_ptr(), nq.qstate.data_ptr(),
                           K, 1 / 0.07, 1.0, qpack.data_ptr(), ds.data_ptr(), None, 1, None, 0, 0, st)
                _cabi.call("mscl_infonce_partial", qpack.data_ptr(), M, nq.queue_tf32.data_ptr(), ds.data_ptr(), K, 0,
                           part.data_ptr(), n_part, 1, st)
                _cabi.call("mscl_infonce_finalize", qpack.data_ptr(), k.data_ptr(), part.data_ptr(), n_part, M, M, 1 / 0.07, 1,
                           row_loss.data_ptr(), dq.data_ptr(), gout.data_ptr(), st)
            for i in range(3):
                whole(i)
            us = time_train(whole, args.iters)
            res["op_us"] = us
            res["op_gbs"] = res["algo_bytes"] / us / 1e3
            res["op_frac"] = res["op_gbs"] / pk
            res["loss"] = float(gout[0, 0])

            # the single-launch form: prep + pass + reduce-add + finalize in one kernel (csrc/infonce_fused.cu)
            n_fp = _cabi.query("mscl_infonce_fused_parts", M, K, sms)
            ws = torch.zeros(16 * M * 4 + 4, device=dev)
            fpart = torch.empty(n_fp, M, fx.PACK_LD, device=dev)
            rowaux = torch.empty(M, 4, device=dev)
            gone = torch.ones(1, device=dev)
            for grad in (1, 0):
                def fused(i, grad=grad):
                    nq = queues[i % n_rot]
                    _cabi.call("mscl_infonce_fused", q.data_ptr(), k.data_ptr(), M, nq.queue_tf32.data_ptr(), nq.birth.data_ptr(),
                               nq.qstate.data_ptr(), K, 1 / 0.07, 1.0, None, 1, ws.data_ptr(), fpart.data_ptr(), n_fp, M, grad, args.flags,
                               row_loss.data_ptr(), rowaux.data_ptr(), gout.data_ptr(), st)
                for i in range(5):
                    fused(i)
                us = time_train(fused, args.iters)
                tag = "fused" if grad else "fused_nograd"
                res[f"{tag}_us"] = us
                res[f"{tag}_frac"] = res["algo_bytes"] / us / 1e3 / pk
            res["fused_loss"] = float(gout[0, 0])

            def fused_fb(i):          # forward launch + the backward kernel that sums the O slabs
                fused(i, 1)
                _cabi.call("mscl_infonce_bwd_slabs", fpart.data_ptr(), n_fp, M, k.data_ptr(), rowaux.data_ptr(), gone.data_ptr(), M,
                           dq.data_ptr(), st)
            for i in range(3):
                fused_fb(i)
            res["fused_fwd_bwd_us"] = time_train(fused_fb, args.iters)
            res["fused_fwd_bwd_frac"] = res["algo_bytes"] / res["fused_fwd_bwd_us"] / 1e3 / pk
            us = res["op_us"]
            rows.append(res)
            print(f"K={K:8d} M={M:4d} parts={n_part:3d}  partial(grad) {res['partial_grad_us']:8.1f} us "
                  f"{res['partial_grad_gbs']:7.0f} GB/s {100 * res['partial_grad_frac']:5.1f}%  {res['partial_grad_tflops']:6.1f} TF | "
                  f"nograd {res['partial_nograd_us']:8.1f} us {100 * res['partial_nograd_frac']:5.1f}% | "
                  f"prep+partial+finalize {us:8.1f} us {100 * res['op_frac']:5.1f}%  loss {res['loss']:.4f} | "
                  f"FUSED op {res['fused_us']:8.1f} us {100 * res['fused_frac']:5.1f}% (nograd {res['fused_nograd_us']:.1f} us, "
                  f"fwd+bwd {res['fused_fwd_bwd_us']:.1f} us {100 * res['fused_fwd_bwd_frac']:.1f}%) "
                  f"loss {res['fused_loss']:.4f}", flush=True)
        del queues
        torch.cuda.empty_cache()
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    with open(args.out, "w") as f:
        json.dump(dict(hbm_peak_gbs=pk, rows=rows), f, indent=1)


if __name__ == "__main__":
    main()
