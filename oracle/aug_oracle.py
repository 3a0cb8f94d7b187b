"""The augmentation front-end written as plain PyTorch ops.  TEST / BENCH INFRASTRUCTURE ONLY.

The product's `SyncMoCoAugmentV5` / `MoCoAugmentV2` / `SyncMoCoAugmentV2` run two CUDA kernels (K8 flow visualiser,
K9 colour pipeline) and refuse host tensors.  This file restates the same pipeline with library ops so that
  (a) tests/test_gpu_kernels.py can check K9 against it for a given draw of the random parameters, and
  (b) bench.py's CPU baseline (`--impl reference`) augments its clips on the host like the reference would
      (kornia ColorJitter(0.4,0.4,0.4,0.1) p=.8 / RandomGrayscale p=.2 / GaussianBlur p=.5 / Normalize,
      common/ssl_aug_v2.py:31-48, common/ssl_aug.py:138-174; flow visualiser common/ssl_aug.py:87-136).
kornia is absent from this image and its random stream is not reproducible, so the parameters
are drawn by the caller (the product's `_color_params`, a handful of torch.rand calls) and passed in.
"""
import math

import torch
import torch.nn.functional as F

from . import mscl_oracle as O


def rgb_to_gray(x):
    return 0.299 * x[:, 0:1] + 0.587 * x[:, 1:2] + 0.114 * x[:, 2:3]


def hue_matrix(h):
    """(n,3,3) RGB->RGB matrices rotating hue by h (fraction of a turn) in YIQ space."""
    theta = h * 2 * math.pi
    c, s = torch.cos(theta), torch.sin(theta)
    yiq = torch.tensor([[0.299, 0.587, 0.114], [0.596, -0.274, -0.322], [0.211, -0.523, 0.312]], device=h.device)
    rot = torch.zeros(h.shape[0], 3, 3, device=h.device)
    rot[:, 0, 0] = 1
    rot[:, 1, 1], rot[:, 1, 2], rot[:, 2, 1], rot[:, 2, 2] = c, -s, s, c
    return torch.linalg.inv(yiq).unsqueeze(0) @ rot @ yiq.unsqueeze(0)


def flip(clips, mask):
    return torch.where(mask.view(-1, 1, 1, 1, 1), torch.flip(clips, [-1]), clips)


def normalize(x, mean, std):
    return (x - mean.to(x.device).view(1, 3, 1, 1, 1)) / std.to(x.device).view(1, 3, 1, 1, 1)


def color_pipeline(x, prm, blur_radius):
    """Brightness -> contrast -> saturation -> hue (applied where prm['jit']), grayscale where prm['gray'],
    separable Gaussian blur with reflect padding where prm['blur']; x (N,3,T,H,W).  The entries of prm have N values
    (one set per clip: the 'params' sync level) or N*T values in (clip, frame) order (one set per frame: the 'batch'
    sync level, whose apply decisions the caller repeats over the frames of a clip); the contrast step is taken about
    the mean luminance of the clip / of the frame respectively."""
    n, _, t = x.shape[:3]
    per_frame = prm["brightness"].numel() == n * t and t > 1
    if per_frame:
        v = lambda p: p.view(n, 1, t, 1, 1)
        mean_dims = (1, 3, 4)
    else:
        v = lambda p: p.view(-1, 1, 1, 1, 1)
        mean_dims = (1, 2, 3, 4)
    y = x * v(prm["brightness"])
    m = rgb_to_gray(y).mean(dim=mean_dims, keepdim=True)
    y = (y - m) * v(prm["contrast"]) + m
    g = rgb_to_gray(y)
    y = (y - g) * v(prm["saturation"]) + g
    hm = hue_matrix(prm["hue"])
    if per_frame:
        y = torch.einsum("ntij,njthw->nithw", hm.view(n, t, 3, 3), y)
    else:
        y = torch.einsum("nij,njthw->nithw", hm, y)
    x = torch.where(v(prm["jit"]), y.clamp(0, 1), x)
    x = torch.where(v(prm["gray"]), rgb_to_gray(x).expand_as(x), x)
    r = blur_radius
    k1 = prm["taps"].to(x.dtype)
    b, c, t, h, w = x.shape
    z = x.reshape(b * c * t, 1, h, w)
    z = F.conv2d(F.pad(z, (r // 2, r // 2, 0, 0), mode="reflect"), k1.view(1, 1, 1, r))
    z = F.conv2d(F.pad(z, (0, 0, r // 2, r // 2), mode="reflect"), k1.view(1, 1, r, 1))
    return torch.where(v(prm["blur"]), z.view(b, c, t, h, w), x)


def augment_view(aug, clips, aux_info, suffix, weak):
    """One view of SyncMoCoAugmentV5.__call__ (common/ssl_aug_v2.py:90-133) on host tensors, with the hyper-parameters
    of the product's aug object `aug` (flip probability, flow suffix, blur radius, mean / std)."""
    mask = torch.rand(clips.shape[0]) < aug.flip_p
    clips = flip(clips, mask)
    if aug.flow_suffix:
        for k in aux_info:
            if k.endswith(aug.flow_suffix + suffix):
                img = flip(O.flow_visualize(aux_info[k]) if aug.visualize else aux_info[k], mask)
                aux_info[k] = normalize(img, aug.mean.view(-1), aug.std.view(-1)) if aug.normalize_flow else img
    if not weak:
        frames = clips.shape[2] if aug.sync_level[0 if suffix == "_q" else 1] == "batch" else 1
        clips = color_pipeline(clips, aug._color_params(clips.shape[0], clips.device, frames), aug.blur_radius)
    return normalize(clips, aug.mean.view(-1), aug.std.view(-1)), aux_info


def augment(aug, im_q, im_k, aux_info):
    im_q, aux_info = augment_view(aug, im_q, aux_info, "_q", aug.weak_aug[0])
    im_k, aux_info = augment_view(aug, im_k, aux_info, "_k", aug.weak_aug[1])
    return im_q, im_k, aux_info
