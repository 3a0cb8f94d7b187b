"""GPU debug: per-CTA phase timeline of the tcgen05 InfoNCE kernel (needs a MSCL_TIMELINE=1 build)."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from mscl_b200 import functional as fx, _cabi

M, K = int(sys.argv[1]) if len(sys.argv) > 1 else 96, int(sys.argv[2]) if len(sys.argv) > 2 else 65536
BURST = int(sys.argv[3]) if len(sys.argv) > 3 else 1      # launches back to back before the timeline is read (clock under load)
g = torch.Generator().manual_seed(0)
q = torch.nn.functional.normalize(torch.randn(M, 128, generator=g), dim=1).cuda()
kp = torch.nn.functional.normalize(torch.randn(M, 128, generator=g), dim=1).cuda()
nq = fx.NegativeQueue(K)
nq.load(torch.nn.functional.normalize(torch.randn(128, K, generator=g), dim=0), torch.ones(K, dtype=torch.long), 0)
flush = torch.empty(64 * 1024 * 1024, device="cuda")
lib = _cabi.load()
for it in range(4):
    flush.fill_(it)
    qd = q.clone().requires_grad_(True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    _cabi.start_timing(["mscl_infonce_partial"])
    for _ in range(BURST):
        out, _ = fx.infonce(qd, kp, nq, M, 0.07)
    rec = _cabi.stop_timing()
    buf = (ctypes.c_ulonglong * (148 * 32))()
    assert lib.mscl_debug_timeline(buf, 148 * 32) == 0
    t = np.array(buf, dtype=np.int64).reshape(148, 32)
    base = t[:, 0].min()
    rel = (t - base) / 1e3
    names = {0: "entry", 1: "setup done", 2: "Q staged", 8: "S0", 9: "S1", 10: "S2", 11: "S3", 12: "S4", 13: "S5", 14: "S6", 20: "P0", 21: "P1", 25: "P5", 26: "P6",
             16: "mma: before issue M1(4)", 17: "mma: M1(4) issued", 28: "mma: Q in TMEM seen", 29: "mma: tile 0 landed", 18: "mma: P3 seen", 19: "mma: M2(3) issued", 23: "P3",
             4: "TMA issued", 5: "softmax done", 6: "O full", 7: "end"}
    print(f"iter {it}: event {rec['mscl_infonce_partial'][-1][0]*1e3:.1f} us; kernel span {(t[:, 7].max() - base)/1e3:.2f} us")
    cyc = (t[:, 31] - t[:, 30]).astype(np.float64)
    ns = (t[:, 3] - t[:, 0]).astype(np.float64)
    print(f"   SM clock during the kernel: {np.median(cyc / ns):.3f} GHz (clock64 / globaltimer, warp 0)")
    for k in sorted(names, key=lambda k: np.median(rel[:, k])):
        print(f"   {names[k]:>24}: min {rel[:, k].min():6.2f}  median {np.median(rel[:, k]):6.2f}  max {rel[:, k].max():6.2f} us")
