"""GPU augmentation front-end under the reference's registry names, without kornia.

Adjacent to the hot path: it has to exist for the config to build
and for (N,2,2T,H,W) flow input to become the 3-channel images the flow encoder eats.
PyTorch ops on the device; the deterministic pieces (flow colour-wheel visualisation,
flip given a mask, normalisation) follow common/ssl_aug.py:87-136 and
common/ssl_aug_v2.py:50-133 exactly, the random colour pipeline reproduces the
reference's distribution (ColorJitter(0.4,0.4,0.4,0.1) p=.8, grayscale p=.2, Gaussian blur
p=.5, decisions shared by the frames of a clip) but not kornia's random stream.
"""
import math

import torch
import torch.nn.functional as F

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
        """flip: optional bool (N,) mask; the colour image of those samples is mirrored along W."""
        if flows.is_cuda:       # K8: lookup + flip in one pass
            from .. import functional as fx
            return fx.flow_visualize(flows.contiguous(), None if flip is None else flip.to(torch.uint8))
        img = self._torch(flows)
        if flip is not None:
            img = torch.where(flip.view(-1, 1, 1, 1, 1), torch.flip(img, [-1]), img)
        return img

    def _torch(self, flows):
        """The reference's op sequence and dtypes (float32 up to f and 1 - f, float64 interpolation)."""
        wheel = self.colorwheel.to(flows.device)
        ncols = wheel.shape[0]
        u, v = flows[:, 0], flows[:, 1]                       # (N,T,H,W)
        rad = torch.sqrt(torch.square(u) + torch.square(v))
        a = torch.atan2(-v, -u) / math.pi
        fk = (a + 1) / 2 * (ncols - 1)
        k0 = torch.floor(fk).long()
        k1 = k0 + 1
        k1[k1 == ncols] = 0
        f = fk - k0                    # float32 (ssl_aug.py:108); (1 - f) below is a float32 op as well
        omf = (1 - f).double()
        f = f.double()
        inside = rad <= 1
        rad_d = rad.double()
        chans = []
        for i in range(3):
            tmp = wheel[:, i]
            col = omf * (tmp[k0] / 255.0) + f * (tmp[k1] / 255.0)
            col = torch.where(inside, 1 - rad_d * (1 - col), col * 0.75)
            chans.append(torch.floor(255 * col).to(torch.uint8).float() / 255)   # uint8 round trip as in the reference
        return torch.stack(chans, dim=1)


@SSL_AUGS.register_module()
class IdentityAug:
    def __init__(self, **kwargs):
        pass

    def __call__(self, clips, im_k=None, aux_info=None):
        if im_k is None and aux_info is None:
            return clips                     # reference signature (common/ssl_aug.py:178-183)
        return clips, im_k, aux_info


def _rgb_to_gray(x):
    return (0.299 * x[:, 0:1] + 0.587 * x[:, 1:2] + 0.114 * x[:, 2:3])


def _hue_matrix(h):
    """(n,3,3) RGB->RGB matrices rotating hue by h (fraction of a turn, per sample) in YIQ space."""
    theta = h * 2 * math.pi
    c, s = torch.cos(theta), torch.sin(theta)
    yiq = torch.tensor([[0.299, 0.587, 0.114], [0.596, -0.274, -0.322], [0.211, -0.523, 0.312]], device=h.device)
    inv = torch.linalg.inv(yiq)
    rot = torch.zeros(h.shape[0], 3, 3, device=h.device)
    rot[:, 0, 0] = 1
    rot[:, 1, 1], rot[:, 1, 2], rot[:, 2, 1], rot[:, 2, 2] = c, -s, s, c
    return inv.unsqueeze(0) @ rot @ yiq.unsqueeze(0)


def _hue_shift(x, h):
    """Rotate hue by h (fraction of a turn, per sample) in YIQ space."""
    return torch.einsum("nij,njthw->nithw", _hue_matrix(h), x)


@SSL_AUGS.register_module()
class SyncMoCoAugmentV5:
    def __init__(self, crop_size, flip_transform=dict(p=0.5, same_on_batch=False), sync_level="batch", t=None,
                 flow_suffix="flow_imgs", img_width=112, visualize=True, weak_aug=(False, False), normalize_flow=False):
        if isinstance(sync_level, str):
            sync_level = (sync_level, sync_level)
        assert all(v in ("batch", "params") for v in sync_level)
        self.mean = torch.tensor([0.485, 0.456, 0.406]).view(1, 3, 1, 1, 1)
        self.std = torch.tensor([0.229, 0.224, 0.225]).view(1, 3, 1, 1, 1)
        self.visualizer = FlowVisualizer() if visualize else (lambda x: x)
        self.flow_suffix = flow_suffix
        self.img_width = img_width
        self.flip_p = flip_transform["p"] if flip_transform else 0.0
        self.weak_aug = weak_aug
        self.normalize_flow = normalize_flow
        self.blur_radius = int(0.1 * crop_size) // 2 * 2 + 1

    def _normalize(self, x):
        return (x - self.mean.to(x.device)) / self.std.to(x.device)

    def flip(self, clips, mask):
        """Mirror the clips selected by the boolean mask along W (deterministic piece)."""
        # torch.where keeps shapes static and needs no host synchronisation (a boolean-index copy does)
        return torch.where(mask.view(-1, 1, 1, 1, 1), torch.flip(clips, [-1]), clips)

    def forward_flip(self, clips, aux_info, suffix="_q", flip_clips=True):
        n = clips.shape[0]
        mask = torch.rand(n, device=clips.device) < self.flip_p
        if flip_clips:
            clips = self.flip(clips, mask)
        if self.flow_suffix:
            full = self.flow_suffix + suffix
            for k in aux_info:
                if k.endswith(full):
                    # only the image is mirrored; the u component keeps its sign (ssl_aug_v2.py:111-117)
                    if isinstance(self.visualizer, FlowVisualizer):
                        img = self.visualizer(aux_info[k], mask)
                    else:
                        img = self.flip(self.visualizer(aux_info[k]), mask)
                    if self.normalize_flow:
                        img = self._normalize(img)
                    aux_info[k] = img
        return clips, aux_info, mask

    def _color_params(self, n, dev):
        """One draw of every decision / parameter of the colour pipeline for n clips ('batch' sync level: shared by
        the frames of a clip): ColorJitter(0.4,0.4,0.4,0.1) p=.8, grayscale p=.2, blur p=.5, one sigma per call."""
        rnd = lambda lo, hi: torch.empty(n, device=dev).uniform_(lo, hi)
        prm = dict(jit=torch.rand(n, device=dev) < 0.8, brightness=rnd(0.6, 1.4), contrast=rnd(0.6, 1.4),
                   saturation=rnd(0.6, 1.4), hue=rnd(-0.1, 0.1), gray=torch.rand(n, device=dev) < 0.2,
                   blur=torch.rand(n, device=dev) < 0.5, sigma=float(torch.empty(1).uniform_(0.1, 2.0)))
        r = self.blur_radius
        ax = torch.arange(r, device=dev, dtype=torch.float32) - r // 2
        k1 = torch.exp(-ax ** 2 / (2 * prm["sigma"] ** 2))
        prm["taps"] = k1 / k1.sum()
        return prm

    def _color_torch(self, x, prm):
        """The colour pipeline as PyTorch ops (host tensors, and the reference the fused kernel is tested against)."""
        v = lambda t: t.view(-1, 1, 1, 1, 1)
        y = x * v(prm["brightness"])                                          # brightness
        m = _rgb_to_gray(y).mean(dim=(1, 2, 3, 4), keepdim=True)
        y = (y - m) * v(prm["contrast"]) + m                                 # contrast
        g = _rgb_to_gray(y)
        y = (y - g) * v(prm["saturation"]) + g                               # saturation
        y = _hue_shift(y, prm["hue"])                                         # hue
        x = torch.where(v(prm["jit"]), y.clamp(0, 1), x)
        x = torch.where(v(prm["gray"]), _rgb_to_gray(x).expand_as(x), x)
        # blur every clip at a fixed shape and select: no data-dependent shapes, no host synchronisation
        r = self.blur_radius
        k1 = prm["taps"].to(x.dtype)
        b, c, t, h, w = x.shape
        z = x.reshape(b * c * t, 1, h, w)
        z = F.conv2d(F.pad(z, (r // 2, r // 2, 0, 0), mode="reflect"), k1.view(1, 1, 1, r))
        z = F.conv2d(F.pad(z, (0, 0, r // 2, r // 2), mode="reflect"), k1.view(1, 1, r, 1))
        return torch.where(v(prm["blur"]), z.view(b, c, t, h, w), x)

    def _color(self, x):
        return self._color_torch(x, self._color_params(x.shape[0], x.device))

    def _pack_params(self, prm, flip, weak):
        """(n,16) float rows for the fused kernel (include/mscl_b200.h, K9)."""
        n = flip.shape[0]
        f32 = lambda t: t.to(torch.float32).view(n, 1)
        if weak:
            z = torch.zeros(n, 15, device=flip.device)
            return torch.cat([f32(flip), z], dim=1).contiguous()
        return torch.cat([f32(flip), f32(prm["jit"]), f32(prm["brightness"]), f32(prm["contrast"]), f32(prm["saturation"]),
                          _hue_matrix(prm["hue"]).reshape(n, 9), f32(prm["gray"]), f32(prm["blur"])], dim=1).contiguous()

    def _view(self, clips, aux_info, suffix, weak):
        """One view: flip decision, flow images, RGB colour pipeline + Normalize."""
        if not clips.is_cuda:
            clips, aux_info, _ = self.forward_flip(clips, aux_info, suffix)
            return self._normalize(clips if weak else self._color(clips)), aux_info
        from .. import functional as fx      # K9: flip + colour + blur + normalise in one pass over the clip
        clips, aux_info, mask = self.forward_flip(clips, aux_info, suffix, flip_clips=False)
        prm = self._color_params(clips.shape[0], clips.device)
        norm = torch.cat([self.mean.view(-1), self.std.view(-1)]).to(clips.device)
        out = fx.color_pipeline(clips.contiguous().float(), self._pack_params(prm, mask, weak), prm["taps"].contiguous(), norm)
        return out, aux_info

    def __call__(self, im_q, im_k, aux_info):
        im_q, aux_info = self._view(im_q, aux_info, "_q", self.weak_aug[0])
        im_k, aux_info = self._view(im_k, aux_info, "_k", self.weak_aug[1])
        return im_q, im_k, aux_info
