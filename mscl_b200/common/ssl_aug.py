"""GPU augmentation front-end under the reference's registry names, without kornia.

Adjacent to the hot path: it has to exist for the config to build
and for (N,2,2T,H,W) flow input to become the 3-channel images the flow encoder eats.
Two kernels do the pixel work (K8 flow_visualize, K9 color_pipeline); this file only draws the
random decisions -- on the HOST (a few hundred numbers per step: drawing and packing them with device ops was ~100 tiny
launches and 1.4 ms of an idle GPU per step, profiles/r02_timeline_host_g1.txt) -- packs them and uploads them in one
asynchronous copy per view.  The deterministic pieces (flow colour-wheel
visualisation, flip given a mask, normalisation) follow common/ssl_aug.py:87-136 and
common/ssl_aug_v2.py:50-133 exactly, the random colour pipeline reproduces the
reference's distribution (ColorJitter(0.4,0.4,0.4,0.1) p=.8, grayscale p=.2, Gaussian blur
p=.5, decisions shared by the frames of a clip) but not kornia's random stream.

CUDA tensors only: a host tensor raises MsclError.  The same pipeline written as PyTorch ops
lives in oracle/aug_oracle.py (test / CPU-baseline infrastructure).
"""
import math

import numpy as np
import torch
import torch.distributed as dist
from .._cabi import MsclError
from ..registry import SSL_AUGS


def make_colorwheel():
    """55-entry Middlebury colour wheel (Baker et al., ICCV 2007) as a float tensor (55, 3):
    six hue segments; inside a segment one channel ramps as floor(255*i/len)."""
    segs = [(15, 0, 1, +1), (6, 1, 0, -1), (4, 1, 2, +1), (11, 2, 1, -1), (13, 2, 0, +1), (6, 0, 2, -1)]
    rows = []
    for n, full, ramp, sign in segs:   # `full` channel saturated, `ramp` channel rising/falling
        for i in range(n):
            rgb = [0.0, 0.0, 0.0]
            rgb[full] = 255.0
            r = math.floor(255 * i / n)
            rgb[ramp] = float(r if sign > 0 else 255 - r)
            rows.append(rgb)
    return torch.tensor(rows, dtype=torch.float64)


class FlowVisualizer:
    """(N,2,T,H,W) flow -> (N,3,T,H,W) colour-wheel image in [0,1] (common/ssl_aug.py:87-136)."""

    def __init__(self):
        self.colorwheel = make_colorwheel()

    def __call__(self, flows, flip=None):
        """flip: optional bool (N,) mask; the colour image of those samples is mirrored along W (K8: lookup +
        flip in one pass)."""
        from .. import functional as fx
        return fx.flow_visualize(flows.contiguous(), None if flip is None else flip.to(torch.uint8))


@SSL_AUGS.register_module()
class IdentityAug:
    def __init__(self, **kwargs):
        pass

    def __call__(self, clips, im_k=None, aux_info=None):
        if im_k is None and aux_info is None:
            return clips                     # reference signature (common/ssl_aug.py:178-183)
        return clips, im_k, aux_info

    def forward_with_flow(self, im_q, im_k, flow_q, flow_k, aux_info):
        """What MoDist.forward_train calls (recognizers/modist.py:79)."""
        return im_q, im_k, flow_q, flow_k, aux_info


_DEV_CONST = {}


def _dev_const(name, device, make):
    """A small constant tensor, built on the host ONCE per device and cached.  Building it per call (`torch.tensor(...,
    device=cuda)`, a pageable `.to(device)`, `torch.linalg.inv` on the device) costs a host <-> device synchronisation each
    time: the augmentation did that six times per step, each one draining the host's run-ahead (profiles/r02_timeline_host_*)."""
    key = (name, str(device))
    t = _DEV_CONST.get(key)
    if t is None:
        t = _DEV_CONST[key] = make().to(device)
    return t


_YIQ = [[0.299, 0.587, 0.114], [0.596, -0.274, -0.322], [0.211, -0.523, 0.312]]


def _hue_matrix(h):
    """(n,3,3) RGB->RGB matrices rotating hue by h (fraction of a turn, per sample) in YIQ space."""
    theta = h * 2 * math.pi
    c, s = torch.cos(theta), torch.sin(theta)
    yiq = _dev_const("yiq", h.device, lambda: torch.tensor(_YIQ))
    inv = _dev_const("yiq_inv", h.device, lambda: torch.linalg.inv(torch.tensor(_YIQ)))
    rot = torch.zeros(h.shape[0], 3, 3, device=h.device)
    rot[:, 0, 0] = 1
    rot[:, 1, 1], rot[:, 1, 2], rot[:, 2, 1], rot[:, 2, 2] = c, -s, s, c
    return inv.unsqueeze(0) @ rot @ yiq.unsqueeze(0)


@SSL_AUGS.register_module()
class SyncMoCoAugmentV5:
    def __init__(self, crop_size, flip_transform=dict(p=0.5, same_on_batch=False), sync_level="batch", t=None,
                 flow_suffix="flow_imgs", img_width=112, visualize=True, weak_aug=(False, False), normalize_flow=False):
        if isinstance(sync_level, str):
            sync_level = (sync_level, sync_level)
        assert all(v in ("batch", "params") for v in sync_level)
        # per view (query, key): 'batch' = toVideoAug (ssl_aug.py:56-60): the frames of a clip share the APPLY decisions
        # only, the jitter factors are drawn per frame; 'params' = toConsistentAug (:62-66): they share the factors too
        self.sync_level = tuple(sync_level)
        self.mean = torch.tensor([0.485, 0.456, 0.406]).view(1, 3, 1, 1, 1)
        self.std = torch.tensor([0.229, 0.224, 0.225]).view(1, 3, 1, 1, 1)
        self.visualize = bool(visualize)
        self.visualizer = FlowVisualizer() if visualize else (lambda x: x)
        self.flow_suffix = flow_suffix
        self.img_width = img_width
        self.flip_p = flip_transform["p"] if flip_transform else 0.0
        self.weak_aug = weak_aug
        self.normalize_flow = normalize_flow
        self.blur_radius = int(0.1 * crop_size) // 2 * 2 + 1

    def _normalize(self, x):
        mean = _dev_const("aug_mean", x.device, lambda: self.mean)
        std = _dev_const("aug_std", x.device, lambda: self.std)
        return (x - mean) / std

    def _norm_table(self, device):
        """[mean3, std3] on `device` (cached: a per-call `.to(device)` of a host tensor synchronises the stream)."""
        return _dev_const("aug_norm", device, lambda: torch.cat([self.mean.view(-1), self.std.view(-1)]))

    def flip(self, clips, mask):
        """Mirror the clips selected by the boolean mask along W (deterministic piece)."""
        # torch.where keeps shapes static and needs no host synchronisation (a boolean-index copy does)
        return torch.where(mask.view(-1, 1, 1, 1, 1), torch.flip(clips, [-1]), clips)

    def _rng(self):
        """The host generator of this object's random decisions: seeded from torch's seed (`torch.manual_seed` makes the
        augmentation reproducible) and the rank, advanced by nothing else (the CPU default generator, which draws the
        shuffle permutations exactly as the reference does, is left alone)."""
        if getattr(self, "_np_rng", None) is None:
            rank = dist.get_rank() if dist.is_available() and dist.is_initialized() else 0
            self._np_rng = np.random.default_rng([int(torch.initial_seed()) % (1 << 32), rank])
        return self._np_rng

    def forward_flip(self, clips, aux_info, suffix="_q", flip_clips=True, mask=None):
        """mask: the flip decisions (bool or uint8 (N,) on the device) when the caller has drawn them already."""
        n = clips.shape[0]
        if mask is None:
            mask = torch.from_numpy(self._rng().random(n) < self.flip_p).to(clips.device)
        if flip_clips:
            clips = self.flip(clips, mask.bool())
        if self.flow_suffix:
            full = self.flow_suffix + suffix
            for k in aux_info:
                if k.endswith(full):
                    # only the image is mirrored; the u component keeps its sign (ssl_aug_v2.py:111-117)
                    if isinstance(self.visualizer, FlowVisualizer):
                        img = self.visualizer(aux_info[k], mask)
                    else:
                        img = self.flip(self.visualizer(aux_info[k]), mask.bool())
                    if self.normalize_flow:
                        img = self._normalize(img)
                    aux_info[k] = img
        return clips, aux_info, mask

    def _color_params(self, n, dev, frames=1):
        """One draw of every decision / parameter of the colour pipeline for n clips: ColorJitter(0.4,0.4,0.4,0.1) p=.8,
        grayscale p=.2, blur p=.5, one sigma per call.  The apply decisions are per clip.  frames = 1: one set of jitter
        factors per clip ('params' sync level); frames = T: one set per frame ('batch' sync level), every tensor then
        has n*T entries in (clip, frame) order, the decisions repeated over the frames of a clip."""
        d = self._draw_host(n, frames)
        return {k: (torch.from_numpy(v).to(dev) if isinstance(v, np.ndarray) else v) for k, v in d.items()}

    def _draw_host(self, n, frames=1):
        """`_color_params` as NumPy arrays on the host."""
        g = self._rng()
        nf = n * frames
        rnd = lambda lo, hi: g.uniform(lo, hi, nf).astype(np.float32)
        dec = lambda p: np.repeat(g.random(n) < p, frames)
        d = dict(jit=dec(0.8), brightness=rnd(0.6, 1.4), contrast=rnd(0.6, 1.4), saturation=rnd(0.6, 1.4),
                 hue=rnd(-0.1, 0.1), gray=dec(0.2), blur=dec(0.5), sigma=float(g.uniform(0.1, 2.0)))
        r = self.blur_radius
        ax = np.arange(r, dtype=np.float32) - r // 2
        k1 = np.exp(-ax ** 2 / np.float32(2 * d["sigma"] ** 2))
        d["taps"] = (k1 / k1.sum()).astype(np.float32)
        return d

    def _pack_host(self, d, flip, weak):
        """`_pack_params` on the host: d from `_draw_host`, flip a bool array with one entry per row."""
        n = flip.shape[0]
        out = np.zeros((n, 16), dtype=np.float32)
        out[:, 0] = flip
        if weak:
            return out
        out[:, 1], out[:, 2], out[:, 3], out[:, 4] = d["jit"], d["brightness"], d["contrast"], d["saturation"]
        theta = d["hue"].astype(np.float64) * 2 * math.pi
        rot = np.zeros((n, 3, 3))
        rot[:, 0, 0] = 1
        rot[:, 1, 1], rot[:, 1, 2], rot[:, 2, 1], rot[:, 2, 2] = np.cos(theta), -np.sin(theta), np.sin(theta), np.cos(theta)
        yiq = np.array(_YIQ, dtype=np.float32).astype(np.float64)
        out[:, 5:14] = (np.linalg.inv(yiq)[None] @ rot @ yiq[None]).reshape(n, 9)
        out[:, 14], out[:, 15] = d["gray"], d["blur"]
        return out

    def _pack_params(self, prm, flip, weak):
        """(n,16) float rows for the fused kernel (include/mscl_b200.h, K9)."""
        n = flip.shape[0]
        f32 = lambda t: t.to(torch.float32).view(n, 1)
        if weak:
            z = torch.zeros(n, 15, device=flip.device)
            return torch.cat([f32(flip), z], dim=1).contiguous()
        return torch.cat([f32(flip), f32(prm["jit"]), f32(prm["brightness"]), f32(prm["contrast"]), f32(prm["saturation"]),
                          _hue_matrix(prm["hue"]).reshape(n, 9), f32(prm["gray"]), f32(prm["blur"])], dim=1).contiguous()

    def _view(self, clips, aux_info, suffix, weak, flow=None):
        """One view: flip decision, flow images, RGB colour pipeline + Normalize.  `flow`: an optional flow clip that
        is mirrored for the same samples (SyncMoCoAugmentV2.forward_with_flow)."""
        if not clips.is_cuda:
            raise MsclError(f"{type(self).__name__} runs on CUDA tensors only (no CPU fallback)")
        from .. import functional as fx      # K9: flip + colour + blur + normalise in one pass over the clip
        n = clips.shape[0]
        frames = clips.shape[2] if self.sync_level[0 if suffix == "_q" else 1] == "batch" else 1
        # every decision and parameter of the view drawn and packed on the host and fetched by ONE small kernel from pinned
        # memory (functional.HostStage: not through the copy engine, where it would wait behind the loader's next batch):
        # [n*frames, 16] K9 rows | the blur taps | the n flip decisions
        flip = self._rng().random(n) < self.flip_p
        d = self._draw_host(n, frames)
        rows = self._pack_host(d, np.repeat(flip, frames), weak)
        table, taps, mask = fx.host_stage(clips.device).upload([rows, d["taps"], flip.astype(np.uint8)], clips.device)
        clips, aux_info, _ = self.forward_flip(clips, aux_info, suffix, flip_clips=False, mask=mask)
        out = fx.color_pipeline(clips.contiguous().float(), table.view(n * frames, 16), taps, self._norm_table(clips.device))
        if flow is not None:
            flow = self.flip(flow, mask.bool())
        return out, aux_info, flow

    def __call__(self, im_q, im_k, aux_info):
        im_q, aux_info, _ = self._view(im_q, aux_info, "_q", self.weak_aug[0])
        im_k, aux_info, _ = self._view(im_k, aux_info, "_k", self.weak_aug[1])
        return im_q, im_k, aux_info


@SSL_AUGS.register_module()
class SyncMoCoAugmentV2(SyncMoCoAugmentV5):
    """Temporally consistent MoCo-v2 augmentation of the `moco_r*_consistent_*` configs (common/ssl_aug.py:249-332):
    per-clip flip, ColorJitter(0.4,0.4,0.4,0.1) p=.8, grayscale p=.2, Gaussian blur p=.5, Normalize -- V5 without the
    flow visualiser and the weak-augmentation switch, plus `forward_with_flow` (MoDist), which mirrors a flow clip for
    the same samples as its RGB clip when `with_flow` is set."""

    def __init__(self, crop_size, flip_transform=dict(p=0.5, same_on_batch=False), sync_level="batch", t=None,
                 with_flow=False, img_width=112):
        assert sync_level in ("batch", "params")      # 'batch' -> toVideoAug, 'params' -> toConsistentAug (ssl_aug.py:259-262)
        super().__init__(crop_size, flip_transform=flip_transform, sync_level=sync_level, t=t, flow_suffix=None,
                         img_width=img_width, visualize=False, weak_aug=(False, False), normalize_flow=False)
        self.with_flow = with_flow

    def forward_with_flow(self, im_q, im_k, flow_q, flow_k, aux_info):
        im_q, aux_info, fq = self._view(im_q, aux_info, "_q", False, flow_q if self.with_flow else None)
        im_k, aux_info, fk = self._view(im_k, aux_info, "_k", False, flow_k if self.with_flow else None)
        return im_q, im_k, (fq if self.with_flow else flow_q), (fk if self.with_flow else flow_k), aux_info


@SSL_AUGS.register_module()
class MoCoAugmentV2(SyncMoCoAugmentV5):
    """Frame-level (temporally inconsistent) MoCo-v2 augmentation of `moco_r18_lr3e-2.py` (common/ssl_aug.py:214-246):
    the clip is treated as N*T independent images -- colour-jitter parameters, grayscale and flip are drawn per FRAME;
    the two `RandomApply` wrappers (jitter p=.8, blur p=.5) draw ONE decision per call for the whole batch.
    Same K9 kernel, launched on the frames as one-frame clips."""

    def __init__(self, crop_size):
        super().__init__(crop_size, flow_suffix=None, visualize=False)

    def single_cal(self, clips):
        if not clips.is_cuda:
            raise MsclError("MoCoAugmentV2 runs on CUDA tensors only (no CPU fallback)")
        from .. import functional as fx
        n, c, t, h, w = clips.shape
        frames = clips.permute(0, 2, 1, 3, 4).reshape(n * t, c, 1, h, w).contiguous().float()
        dev = clips.device
        prm = self._color_params(n * t, dev)
        prm["jit"] = (torch.rand(1, device=dev) < 0.8).expand(n * t)
        prm["blur"] = (torch.rand(1, device=dev) < 0.5).expand(n * t)
        mask = torch.rand(n * t, device=dev) < 0.5
        norm = self._norm_table(dev)
        out = fx.color_pipeline(frames, self._pack_params(prm, mask, False), prm["taps"].contiguous(), norm)
        return out.view(n, t, c, h, w).permute(0, 2, 1, 3, 4).contiguous()

    def __call__(self, im_q, im_k, aux_info):
        return self.single_cal(im_q), self.single_cal(im_k), aux_info
