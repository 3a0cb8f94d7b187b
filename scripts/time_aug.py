"""GPU: how long the augmentation front-end and the optimizer tail of the bench step take (CUDA events)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import mscl_b200
from mscl_b200 import functional as fx
from mscl_b200.configs import mscl_r18_model
from bench import make_host_batch

dev = torch.device("cuda", 0)
model = mscl_b200.build_model(mscl_r18_model(K=65536)).to(dev)
b = {k: v.to(dev) for k, v in make_host_batch(32, 1, False).items()}
tab = fx.fra_table(device=dev)


def timeit(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


fq = fx.fra(b["flow_q"], b["cid_q"], tab, "planar")
fk = fx.fra(b["flow_k"], b["cid_k"], tab, "planar")
aug = model.aug_gpu
print("aug total        %.3f ms" % timeit(lambda: aug(b["imgs_q"], b["imgs_k"], dict(flow_imgs_q=fq, flow_imgs_k=fk))))
print("visualizer x1    %.3f ms" % timeit(lambda: aug.visualizer(fq)))
print("color x1         %.3f ms" % timeit(lambda: aug._color(b["imgs_q"])))
print("normalize x1     %.3f ms" % timeit(lambda: aug._normalize(b["imgs_q"])))
mask = torch.rand(32, device=dev) < 0.5
print("flip x1          %.3f ms" % timeit(lambda: aug.flip(b["imgs_q"], mask)))
params = [p for p in model.parameters() if p.requires_grad]
for p in params:
    p.grad = torch.randn_like(p)
opt = torch.optim.SGD(params, lr=0.02, momentum=0.9, weight_decay=1e-4)
print("clip_grad_norm   %.3f ms" % timeit(lambda: torch.nn.utils.clip_grad_norm_(params, 40.0)))
print("sgd step         %.3f ms" % timeit(lambda: opt.step()))
print("zero_grad(none)+randn skipped; n params %d, %d tensors" % (sum(p.numel() for p in params), len(params)))
