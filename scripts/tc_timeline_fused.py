"""GPU debug: per-CTA phase timeline of the single-launch InfoNCE kernel (needs a MSCL_TIMELINE=1 build)."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from mscl_b200 import functional as fx, _cabi

M, K = int(sys.argv[1]) if len(sys.argv) > 1 else 96, int(sys.argv[2]) if len(sys.argv) > 2 else 65536
FLUSH = int(sys.argv[3]) if len(sys.argv) > 3 else 1      # 1: dirty the L2 before each launch, 0: back-to-back launches, 2: one launch after a device sync
STEP = len(sys.argv) > 4 and sys.argv[4] == "step"       # the step's launch: 3n rows over one queue + (n | 3n) rows, epoch split, over another (n = M / 3)
g = torch.Generator().manual_seed(0)
q = torch.nn.functional.normalize(torch.randn(M, 128, generator=g), dim=1).cuda()
kp = torch.nn.functional.normalize(torch.randn(M, 128, generator=g), dim=1).cuda()
nqs = []
for s in range(6):
    nq = fx.NegativeQueue(K)
    nq.load(torch.nn.functional.normalize(torch.randn(128, K, generator=g), dim=0), torch.ones(K, dtype=torch.long), 0)
    nqs.append(nq)
flush = torch.empty(64 * 1024 * 1024, device="cuda")
lib = _cabi.load()
for it in range(4):
    if FLUSH == 1:
        flush.fill_(it)
    if FLUSH == 2:
        torch.cuda.synchronize()
    entry = "mscl_infonce_fused_multi_x" if STEP else "mscl_infonce_fused"
    _cabi.start_timing([entry])
    for j in range(1 if FLUSH else (3 if STEP else 6)):
        nqs[j]._fresh = False
        qd = q.clone().requires_grad_(True)
        if STEP:
            n = M // 3
            nqs[j + 3]._fresh = False
            q2 = torch.cat([q, q[:n]]).clone().requires_grad_(True)
            k2 = torch.cat([kp, kp[:n]])
            dup = torch.full((4 * n,), -1, dtype=torch.int32)
            dup[2 * n:3 * n] = torch.arange(n, dtype=torch.int32) + nqs[j + 3].ptr
            ow = nqs[j + 3].enqueue(k2[2 * n:3 * n].contiguous(), save=True)
            nqs[j + 3]._fresh = False
            fx.infonce_multi([dict(q=qd, kpos=kp, nq=nqs[j], rows_per_group=n, T=0.07),
                              dict(q=q2, kpos=k2, nq=nqs[j + 3], rows_per_group=n, T=0.07, dup_slot=dup.cuda(), dup_age=1,
                                   overwritten=ow, row_split=n)])
        else:
            out, _ = fx.infonce(qd, kp, nqs[j], M, 0.07)
    rec = _cabi.stop_timing()
    buf = (ctypes.c_ulonglong * (148 * 32))()
    assert lib.mscl_debug_timeline_fused(buf, 148 * 32) == 0
    t = np.array(buf, dtype=np.int64).reshape(148, 32)
    base = t[:, 0].min()
    rel = (t - base) / 1e3
    names = {0: "entry", 1: "setup done", 2: "Q staged", 8: "S0", 9: "S1", 10: "S2", 11: "S3", 12: "S4", 20: "P0", 21: "P1", 22: "P2", 23: "P3",
             24: "P4", 4: "TMA issued", 5: "softmax done", 6: "O full", 7: "ticket taken", 16: "epilogue done", 17: "finalize done", 3: "exit",
             12: "positives landed (thread 64)", 13: "row info reduced (thread 64)", 14: "softmax warps enter the tile loop", 15: "MMA1(0) complete (s_full[0] fires)", 18: "producer: dependency wait over", 19: "scales of tile 0 ready", 28: "mma: Q in TMEM seen", 29: "mma: tile 0 landed, MMA1(0) issued",
             30: "mma: tile 1 landed, MMA1(1) issued", 31: "mma: tile 2 landed, MMA1(2) issued",
             24: "last CTA: row statistics done (thread 0)", 27: "last CTA: idle warps woke up (thread 0)", 25: "stats warp: REDs performed (fence done)", 26: "stats warp: ticket atomic returned"}
    if STEP:      # 7 steps per CTA: the P4.. stamps share slots 24.. with the stats warp's
        for k in (24, 25, 26, 27):
            names.pop(k)
    print(f"iter {it}: event {rec[entry][-1][0]*1e3:.1f} us; kernel span {(t[:, 3].max() - base)/1e3:.2f} us")
    for k in sorted(names, key=lambda k: np.median(rel[:, k])):
        col = rel[:, k][t[:, k] > 0]
        if col.size:
            print(f"   {names[k]:>36}: min {col.min():6.2f}  median {np.median(col):6.2f}  max {col.max():6.2f} us  ({col.size} CTAs)")
