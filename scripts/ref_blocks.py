"""Per-block timing of the REFERENCE's operation sequence (through the oracle's restatement), BASELINE.md section 3:

    python scripts/ref_blocks.py --device cpu     # host cores of the box (torch.set_num_threads(os.cpu_count()))
    python scripts/ref_blocks.py --device cuda    # the same eager PyTorch ops on one B200: the GPU-vs-GPU bar

Blocks = the rows of SURVEY.md section 8a at the r18 config's sizes (N = 32, C = 128, K = 65536, t = 4):
  a1+a2  decayed snapshot + logits + CE + host top-k (fwd + bwd)           moco.py:481-498, moco_head.py:38-77
  a3     the two cross-modal terms                                          moco_head_v2.py:38-100
  a4     LMCL on (32,128,4,28,28) + 2 x (32,128,4,7,7) (fwd + bwd)          local_cl_head.py:41-73
  a5     momentum EMA over the RGB key side (36.7 M parameters, 88 tensors) moco.py:408-421
  a6     dequeue / enqueue                                                  moco.py:423-440
  a9     FRA on 32 clips of 8 frames (NumPy; CPU only)                      transforms_motion.py:7-29,103-142
TEST / BENCH INFRASTRUCTURE: imports oracle/, never imported by the product.  One JSON line per block.
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import inputs, mscl_oracle as O  # noqa: E402


def timed(fn, dev, warmup=3, iters=20):
    def sync():
        if dev.type == "cuda":
            torch.cuda.synchronize()
    for _ in range(warmup):
        fn()
    ts = []
    for _ in range(iters):
        sync()
        t0 = time.perf_counter()
        fn()
        sync()
        ts.append(time.perf_counter() - t0)
    return float(np.median(ts)) * 1e3


def ce_topk(logits):
    """MoCoHead.loss: CE with target 0 + top-1/5 through a device->host copy and NumPy argsort (accuracy.py:130-149)."""
    labels = torch.zeros(logits.shape[0], dtype=torch.long, device=logits.device)
    acc = O.top_k_accuracy(logits.detach().cpu().numpy(), labels.cpu().numpy(), (1, 5))
    return F.cross_entropy(logits, labels, ignore_index=-1), acc


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--device", default="cpu", choices=["cpu", "cuda"])
    ap.add_argument("--N", type=int, default=32)
    ap.add_argument("--K", type=int, default=65536)
    ap.add_argument("--iters", type=int, default=20)
    args = ap.parse_args()
    dev = torch.device(args.device)
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    N, K, t = args.N, args.K, 4
    inp = inputs.head_inputs(seed=0, N=N, K=K, t=t, hw_rgb=28, hw_flow=7, b_all=N)
    x = {k: (v.to(dev) if isinstance(v, torch.Tensor) else v) for k, v in inp.items()}
    T = 0.07
    meta = dict(device=args.device, cores=threads, N=N, K=K, torch=torch.__version__,
                gpu=torch.cuda.get_device_name(0) if dev.type == "cuda" else None)

    def emit(block, ms, what):
        print(json.dumps(dict(block=block, ms=round(ms, 4), what=what, **meta)), flush=True)

    # a1 + a2
    q = x["q"].clone().requires_grad_(True)

    def a12():
        q.grad = None
        w = O.decayed_weight(x["queue_rgb"], x["count"])
        loss, _ = ce_topk(O.infonce_logits(q, x["k"], w, T))
        loss.backward()
    emit("a1+a2", timed(a12, dev, iters=args.iters), "decayed snapshot + logits + CE + host top-k, fwd+bwd (x7 per step in the reference)")

    # a3
    qf = x["q_f"].clone().requires_grad_(True)
    w_rgb, w_flow = O.decayed_weight(x["queue_rgb"], x["count"]), O.decayed_weight(x["queue_flow"], x["count"])

    def a3():
        q.grad = qf.grad = None
        l1, _ = ce_topk(O.infonce_logits(q, x["k_f"], w_flow, T))
        l2, _ = ce_topk(O.infonce_logits(qf, x["k"], w_rgb, T))
        (l1 + l2).backward()
    emit("a3", timed(a3, dev, iters=args.iters), "rf + fr cross-modal terms given the snapshots, fwd+bwd (x2 per step)")

    # a4
    maps = [x[n].clone().requires_grad_(True) for n in ("q_map", "qf_map", "qaf_map")]

    def a4():
        for m in maps:
            m.grad = None
        x_q = maps[0].mean(dim=(-2, -1))
        x_f = torch.cat((maps[1], maps[2]), dim=2).mean(dim=(-2, -1))
        sim = torch.bmm(F.normalize(x_q, dim=1).transpose(1, 2), F.normalize(x_f, dim=1))
        scores = sim.flatten(0, 1) / T
        labels = torch.arange(t, device=dev).unsqueeze(0).repeat((N, 1)).flatten(0, 1)
        O.top_k_accuracy(scores.detach().cpu().numpy(), labels.cpu().numpy(), (1, 5))
        F.cross_entropy(scores, labels, ignore_index=-1).backward()
    emit("a4", timed(a4, dev, iters=args.iters), "LMCL: pooling + normalise + bmm + CE + host top-k, fwd+bwd")

    # a5: 36,707,392 parameters in 88 tensors, sizes drawn to that total
    rng = np.random.RandomState(0)
    sizes = rng.multinomial(36_707_392 - 88 * 64, np.ones(88) / 88) + 64
    pk = [torch.randn(int(s), device=dev) for s in sizes]
    pq = [torch.randn(int(s), device=dev) for s in sizes]
    m = O.momentum(500, 1000, 0.994)

    def a5():
        for i, new in enumerate(O.ema_update(pk, pq, m)):
            pk[i] = new
    emit("a5", timed(a5, dev, iters=args.iters), "momentum EMA, 36.7 M parameters / 88 tensors (x3 per step)")

    # a6
    queue, count = x["queue_rgb"].clone(), x["count"].clone()
    state = dict(ptr=int(inp["ptr"]))

    def a6():
        state["ptr"] = O.enqueue(queue, count, state["ptr"], x["k"])
    emit("a6", timed(a6, dev, iters=args.iters), "count += 1, block write, ptr (x2 per step)")

    # a9 (NumPy, as in the data-loader workers)
    if dev.type == "cpu":
        clips = [inputs.flow_clip(seed=s, T=8, H=112, W=112) for s in range(N)]

        def a9():
            for c, clip in enumerate(clips):
                O.fra(clip, c % 8)
        emit("a9", timed(a9, dev, warmup=1, iters=3), f"FRA on {N} clips x 8 frames of 112x112 flow (NumPy)")


if __name__ == "__main__":
    main()
