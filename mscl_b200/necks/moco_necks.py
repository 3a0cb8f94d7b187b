"""Necks of the MoCo / MSCL recognizers (PyTorch / cuDNN; not on the accelerated path, but
their parameters are inputs of the EMA kernel and their names are part of the state_dict).

BaseMoCo : global average pool of the last level (necks/base.py:10-24).
TPNMoCo  : the same embedding, plus a 3-level pyramid = FPN with (1,3,3) convs
           (necks/fpn.py:146-227) followed by SEPC pyramid convolutions
           (necks/sepc.py:17-135) -- (necks/base.py:137-175, fpn_video.py:43-136).
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

from ..registry import NECKS


def _trilinear(x, size):
    """F.interpolate(x, size, mode="trilinear") (necks/sepc.py:126-130) as the K7 kernel (PyTorch's
    upsample_trilinear3d runs at ~30 GB/s on these shapes).  fp32 CUDA tensors only: anything else raises MsclError.
    The oracle, which re-uses these module classes on the host, replaces the `upsample` attribute of ITS copies
    (oracle/step.py)."""
    from .. import functional as fx
    return fx.upsample_trilinear(x, size)


class _Conv(nn.Module):
    """Plain conv held under `.conv`, the attribute name mmcv's ConvModule uses."""

    def __init__(self, cin, cout, kernel, padding=0, stride=1):
        super().__init__()
        self.conv = nn.Conv3d(cin, cout, kernel, stride=stride, padding=padding)

    def forward(self, x):
        return self.conv(x)


class _FPN3d(nn.Module):
    def __init__(self, in_channels, out_channels, kernel=(1, 3, 3)):
        super().__init__()
        pad = tuple((k - 1) // 2 for k in kernel)
        self.lateral_convs = nn.ModuleList(_Conv(c, out_channels, 1) for c in in_channels)
        self.fpn_convs = nn.ModuleList(_Conv(out_channels, out_channels, kernel, pad) for _ in in_channels)

    def forward(self, feats):
        lat = [conv(f) for conv, f in zip(self.lateral_convs, feats)]
        for i in range(len(lat) - 1, 0, -1):      # top-down, nearest upsampling
            lat[i - 1] = lat[i - 1] + F.interpolate(lat[i], size=lat[i - 1].shape[2:], mode="nearest")
        return [conv(x) for conv, x in zip(self.fpn_convs, lat)]


class _PConv3D(nn.Module):
    """One pyramid convolution: each level sums a conv of itself, a strided conv of the
    finer level and an upsampled conv of the coarser level (necks/sepc.py:117-135)."""

    def __init__(self, cin, cout, stride):
        super().__init__()
        self.Pconv = nn.ModuleList([nn.Conv3d(cin, cout, 3, padding=1), nn.Conv3d(cin, cout, 3, padding=1),
                                    nn.Conv3d(cin, cout, 3, padding=1, stride=stride)])
        self.relu = nn.ReLU()
        self.upsample = _trilinear

    def forward(self, xs):
        out = []
        for lvl, f in enumerate(xs):
            y = self.Pconv[1](f)
            if lvl > 0:
                y = y + self.Pconv[2](xs[lvl - 1])
            if lvl < len(xs) - 1:
                y = y + self.upsample(self.Pconv[0](xs[lvl + 1]), list(y.shape[2:]))
            out.append(self.relu(y))
        return out


class _SEPC(nn.Module):
    def __init__(self, in_channels, out_channels, stride=(2, 1, 1), iBN=False, Pconv_num=2):
        super().__init__()
        if iBN:
            raise NotImplementedError("iBN=True is not used by the MSCL configs")
        self.Pconvs = nn.ModuleList(_PConv3D(in_channels[i], out_channels, stride) for i in range(Pconv_num))

    def forward(self, xs):
        for p in self.Pconvs:
            xs = p(xs)
        return xs


class _TPNSingle(nn.Module):
    def __init__(self, in_channels, out_channels, fpn_cfg, temporal_modulation_cfg, sepc_cfg, reverse_st):
        super().__init__()
        if temporal_modulation_cfg is not None or reverse_st:
            raise NotImplementedError("temporal modulation / reverse_st are not used by the MSCL configs")
        self.num_tpn_stages = len(in_channels)
        self.fpn = _FPN3d(in_channels, out_channels, tuple(fpn_cfg.get("fpn_kerne_size", (1, 3, 3))))
        self.sepc = _SEPC(**sepc_cfg) if sepc_cfg is not None else None

    def forward(self, x):
        outs = self.fpn(list(x[-self.num_tpn_stages:]))
        return self.sepc(outs) if self.sepc is not None else outs


@NECKS.register_module()
class BaseMoCo(nn.Module):
    def __init__(self):
        super().__init__()
        self.tofc = nn.Sequential(nn.AdaptiveAvgPool3d((1, 1, 1)), nn.Flatten(1))

    def forward(self, x, target=None, pyramid=True):
        return (self.tofc(x[-1]), x), dict()

    def init_weights(self):
        pass


@NECKS.register_module()
class TPNMoCo(nn.Module):
    def __init__(self, in_channels, out_channels,
                 fpn_cfg=dict(fpn_kerne_size=(1, 3, 3), conv_cfg=dict(type="Conv3d")),
                 temporal_modulation_cfg=None, sepc_cfg=None, reverse_st=False, emb_from_bkb=True):
        super().__init__()
        self.tpn = _TPNSingle(in_channels, out_channels, fpn_cfg, temporal_modulation_cfg, sepc_cfg, reverse_st)
        self.tofc = nn.Sequential(nn.AdaptiveAvgPool3d((1, 1, 1)), nn.Flatten(1))
        self.emb_from_bkb = emb_from_bkb
        self.init_weights()

    def init_weights(self):
        # net effect of the reference's nested init calls: xavier-uniform weights, zero biases
        for m in self.tpn.modules():
            if isinstance(m, nn.Conv3d):
                nn.init.xavier_uniform_(m.weight)
                if m.bias is not None:
                    nn.init.constant_(m.bias, 0)

    def forward(self, x, target=None, pyramid=True):
        """pyramid=False: the caller consumes only the embedding (the key side of MSCLWithAug never reads k_mlvl,
        SURVEY.md section 2.4); with emb_from_bkb the embedding does not depend on the pyramid, so it is not built."""
        if self.emb_from_bkb:
            emb = self.tofc(x[-1])
            x = self.tpn(x) if pyramid else list(x)
        else:
            x = self.tpn(x)
            emb = self.tofc(x[-1])
        return (emb, x), {}
