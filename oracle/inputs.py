"""Seeded synthetic inputs shared by the golden generator, the tests and bench.py.

TEST / BENCH INFRASTRUCTURE.  Everything is drawn from a private torch.Generator on
the CPU so the default generator (which drives the shuffle-BN permutation,
moco.py:160) is never disturbed.  Shapes follow SURVEY.md section 8d.
"""
import hashlib

import numpy as np
import torch
import torch.nn.functional as F


def _gen(seed):
    g = torch.Generator(device="cpu")
    g.manual_seed(int(seed))
    return g


def steady_state_count(K, b_all, ptr):
    """Ages of a queue that has been cycling with batch b_all: the block written last has
    age 1 (moco.py:427,437); count[j] = 1 + floor(((ptr-1-j) mod K) / b_all)."""
    j = torch.arange(K, dtype=torch.long)
    return 1 + ((ptr - 1 - j) % K) // b_all


def head_inputs(seed=0, N=8, C=128, K=4096, t=8, hw_rgb=28, hw_flow=7, b_all=None, ptr_blocks=5):
    """Config-1 style inputs: six unit-norm (N,C) feature sets, two unit-norm-column
    (C,K) queues with steady-state ages, and the three LMCL feature maps."""
    g = _gen(seed)
    b_all = b_all or N
    out = {}
    # a shared latent per clip plus row-dependent noise, so that the positives land on
    # a spread of ranks (top-1, top-5, neither) instead of all being lost among K negatives
    z = torch.randn(N, C, generator=g)
    noise = torch.linspace(0.15, 1.6, N).unsqueeze(1)
    for name in ("q", "k", "q_f", "k_f", "q_af", "k_af"):
        out[name] = F.normalize(z + noise * torch.randn(N, C, generator=g), dim=1)
    for name in ("queue_rgb", "queue_flow"):
        out[name] = F.normalize(torch.randn(C, K, generator=g), dim=0)
    ptr = (ptr_blocks * b_all) % K
    out["ptr"] = ptr
    out["count"] = steady_state_count(K, b_all, ptr)
    zt = torch.randn(N, C, t, 1, 1, generator=g) * 0.08 * torch.linspace(0.2, 1.5, N).view(N, 1, 1, 1, 1)
    out["q_map"] = torch.randn(N, C, t, hw_rgb, hw_rgb, generator=g) + zt
    out["qf_map"] = torch.randn(N, C, t, hw_flow, hw_flow, generator=g) + zt
    out["qaf_map"] = torch.randn(N, C, t, hw_flow, hw_flow, generator=g) + 0.5 * zt
    return out


def sibling_head_cases():
    """(name, reference class name, ctor kwargs, rgb level shapes, flow level shapes, uses FRA flow) -- shared with the
    GPU parity test, which rebuilds the same inputs from the seed."""
    return [
        ("pos_head", "MSCLWithAugPosHead", dict(bkb_channels=(48, 24), t=4, T=0.07, mlvl_ids=(0, -1)),
         [(3, 48, 4, 6, 6)], [(3, 24, 4, 3, 3)], True),
        ("pos_head_identity_rgb", "MSCLWithAugPosHead", dict(bkb_channels=(None, 24), t=4, T=0.07, mlvl_ids=(0, -1)),
         [(3, 128, 4, 5, 5)], [(3, 24, 4, 3, 3)], True),
        ("modist_pos_head", "MoDistv2PosHead", dict(bkb_channels=(None, 32), t=8, T=0.1, mlvl_ids=(0, -1)),
         [(2, 128, 8, 7, 7)], [(2, 32, 8, 4, 4)], False),
        ("mlvl_pos_head", "MlvlMSCLWithAugPosHead", dict(bkb_channels=(None, 16), t=4, T=0.07, mlvl_ids=(0, 1, 2),
                                                         mlvl_flow_ids=(-1, -1, -1)),
         [(3, 128, 4, 8, 8), (3, 128, 4, 4, 4), (3, 128, 4, 2, 2)], [(3, 16, 4, 3, 3)], True),
    ]


def sibling_head_inputs(case, seed=7):
    name, _, kw, rgb_shapes, flow_shapes, with_aug = case
    g = torch.Generator().manual_seed(seed + len(name))
    t = kw["t"]
    zt = [torch.randn(rgb_shapes[0][0], 1, t, 1, 1, generator=g) for _ in range(2)]
    q_mlvl = [torch.randn(*s, generator=g) * 0.3 + zt[0] for s in rgb_shapes]
    qf_mlvl = [torch.randn(*s, generator=g) * 0.3 + zt[0] + 0.3 * zt[1] for s in flow_shapes]
    qaf_mlvl = [torch.randn(*s, generator=g) * 0.3 + zt[1] for s in flow_shapes] if with_aug else None
    return q_mlvl, qf_mlvl, qaf_mlvl


def flow_clip(seed=0, T=8, H=112, W=112):
    """One clip of raw optical flow, list of T float32 (H, W, 2) frames."""
    g = _gen(seed)
    x = torch.randn(T, H, W, 2, generator=g) * 3.0
    return [x[i].numpy().copy() for i in range(T)]


def digest(*tensors):
    """Order-sensitive sha1 of raw bytes; detects RNG drift between torch builds."""
    h = hashlib.sha1()
    for t in tensors:
        a = t.detach().cpu().numpy() if isinstance(t, torch.Tensor) else np.asarray(t)
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()
