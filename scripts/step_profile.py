"""GPU: where one bench step spends its device time (torch.profiler, no replay): top kernels by total time."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
import mscl_b200
from mscl_b200 import functional as fx
from mscl_b200.configs import mscl_r18_model
from bench import make_host_batch

cl = "--nchw" not in sys.argv
dev = torch.device("cuda", 0)
torch.backends.cudnn.benchmark = True
torch.manual_seed(0)
model = mscl_b200.build_model(mscl_r18_model(K=65536)).to(dev)
if cl:
    model = model.to(memory_format=torch.channels_last_3d)
model.train()
params = [p for p in model.parameters() if p.requires_grad]
opt = torch.optim.SGD(params, lr=0.02, momentum=0.9, weight_decay=1e-4)
tab = fx.fra_table(device=dev)
bs = [{k: v.to(dev) for k, v in make_host_batch(32, i, False).items()} for i in range(2)]


def step(b):
    aux = dict(flow_imgs_q=fx.fra(b["flow_q"], b["cid_q"], tab, "planar"), flow_imgs_k=fx.fra(b["flow_k"], b["cid_k"], tab, "planar"))
    losses = model(b["imgs_q"], b["imgs_k"], aux, return_loss=True)
    loss, lv = model._parse_losses(losses)
    opt.zero_grad(set_to_none=True)
    loss.backward()
    torch.nn.utils.clip_grad_norm_(params, 40.0)
    opt.step()


for i in range(4):
    step(bs[i % 2])
torch.cuda.synchronize()
N = 3
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for i in range(N):
        step(bs[i % 2])
    torch.cuda.synchronize()
ev = prof.key_averages()
rows = sorted(ev, key=lambda e: -e.device_time_total)
tot = sum(e.device_time_total for e in rows)
print(f"# channels_last={cl}: {tot / N / 1e3:.2f} ms of kernel time per step, {sum(e.count for e in rows) / N:.0f} launches per step")
for e in rows[:45]:
    print(f"{e.device_time_total / N / 1e3:8.3f} ms {100 * e.device_time_total / tot:5.1f}% {e.count / N:7.1f}x  {e.key[:150]}")
