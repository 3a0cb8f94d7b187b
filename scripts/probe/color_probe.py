import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from mscl_b200 import functional as fx
from mscl_b200.common.ssl_aug import SyncMoCoAugmentV5
dev = torch.device("cuda", 0)
aug = SyncMoCoAugmentV5(crop_size=112, sync_level=("batch", "batch"), t=(8, 8), flow_suffix="flow_imgs")
N = 32
x = torch.rand(N, 3, 8, 112, 112, device=dev)
torch.manual_seed(0)
prm = aug._color_params(N, dev)
norm = torch.cat([aug.mean.view(-1), aug.std.view(-1)]).to(dev)
flip = (torch.arange(N, device=dev) % 2).bool()
for blur in (True, False, True, False):
    prm["blur"] = torch.full((N,), blur, device=dev)
    y = fx.color_pipeline(x, aug._pack_params(prm, flip, False), prm["taps"].contiguous(), norm)
fl = torch.randn(N, 2, 16, 112, 112, device=dev)
for _ in range(2):
    fx.flow_visualize(fl, flip.to(torch.uint8))
torch.cuda.synchronize()
