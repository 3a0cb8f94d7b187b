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

    def __call__(self, flows):
        wheel = self.colorwheel.to(flows.device)
        ncols = wheel.shape[0]
        u, v = flows[:, 0], flows[:, 1]                       # (N,T,H,W)
        rad = torch.sqrt(torch.square(u) + torch.square(v))
        a = torch.atan2(-v, -u) / math.pi
        fk = (a + 1) / 2 * (ncols - 1)
        k0 = torch.floor(fk).long()
        k1 = k0 + 1
        k1[k1 == ncols] = 0
        f = (fk - k0).double()
        inside = rad <= 1
        rad_d = rad.double()
        chans = []
        for i in range(3):
            tmp = wheel[:, i]
            col = (1 - f) * (tmp[k0] / 255.0) + f * (tmp[k1] / 255.0)
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


def _hue_shift(x, h):
    """Rotate hue by h (fraction of a turn, per sample) in YIQ space."""
    theta = h * 2 * math.pi
    c, s = torch.cos(theta), torch.sin(theta)
    yiq = torch.tensor([[0.299, 0.587, 0.114], [0.596, -0.274, -0.322], [0.211, -0.523, 0.312]], device=x.device)
    inv = torch.linalg.inv(yiq)
    rot = torch.zeros(x.shape[0], 3, 3, device=x.device)
    rot[:, 0, 0] = 1
    rot[:, 1, 1], rot[:, 1, 2], rot[:, 2, 1], rot[:, 2, 2] = c, -s, s, c
    m = inv.unsqueeze(0) @ rot @ yiq.unsqueeze(0)
    return torch.einsum("nij,njthw->nithw", m, x)


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

    def forward_flip(self, clips, aux_info, suffix="_q"):
        n = clips.shape[0]
        mask = torch.rand(n, device=clips.device) < self.flip_p
        clips = self.flip(clips, mask)
        if self.flow_suffix:
            full = self.flow_suffix + suffix
            for k in aux_info:
                if k.endswith(full):
                    img = self.visualizer(aux_info[k])
                    if self.normalize_flow:
                        img = self._normalize(img)
                    # only the image is mirrored; the u component keeps its sign (ssl_aug_v2.py:111-117)
                    aux_info[k] = self.flip(img, mask)
        return clips, aux_info

    def _color(self, x):
        n, dev = x.shape[0], x.device
        rnd = lambda lo, hi: torch.empty(n, 1, 1, 1, 1, device=dev).uniform_(lo, hi)
        jit = (torch.rand(n, device=dev) < 0.8).view(n, 1, 1, 1, 1)
        y = x * rnd(0.6, 1.4)                                                 # brightness
        m = _rgb_to_gray(y).mean(dim=(1, 2, 3, 4), keepdim=True)
        y = (y - m) * rnd(0.6, 1.4) + m                                      # contrast
        g = _rgb_to_gray(y)
        y = (y - g) * rnd(0.6, 1.4) + g                                      # saturation
        y = _hue_shift(y, torch.empty(n, device=dev).uniform_(-0.1, 0.1))     # hue
        x = torch.where(jit, y.clamp(0, 1), x)
        gray = (torch.rand(n, device=dev) < 0.2).view(n, 1, 1, 1, 1)
        x = torch.where(gray, _rgb_to_gray(x).expand_as(x), x)
        # Gaussian blur p=0.5, one sigma draw per call: blur every clip at a fixed shape and select,
        # so the step has no data-dependent shapes and no host synchronisation
        blur = (torch.rand(n, device=dev) < 0.5).view(n, 1, 1, 1, 1)
        sigma = float(torch.empty(1).uniform_(0.1, 2.0))
        r = self.blur_radius
        ax = torch.arange(r, device=dev, dtype=x.dtype) - r // 2
        k1 = torch.exp(-ax ** 2 / (2 * sigma ** 2))
        k1 = k1 / k1.sum()
        b, c, t, h, w = x.shape
        z = x.reshape(b * c * t, 1, h, w)
        z = F.conv2d(F.pad(z, (r // 2, r // 2, 0, 0), mode="reflect"), k1.view(1, 1, 1, r))
        z = F.conv2d(F.pad(z, (0, 0, r // 2, r // 2), mode="reflect"), k1.view(1, 1, r, 1))
        return torch.where(blur, z.view(b, c, t, h, w), x)

    def __call__(self, im_q, im_k, aux_info):
        im_q, aux_info = self.forward_flip(im_q, aux_info, suffix="_q")
        im_q = self._normalize(im_q if self.weak_aug[0] else self._color(im_q))
        im_k, aux_info = self.forward_flip(im_k, aux_info, suffix="_k")
        im_k = self._normalize(im_k if self.weak_aug[1] else self._color(im_k))
        return im_q, im_k, aux_info
