"""Video ResNets used by the MSCL configs (they stay in PyTorch / cuDNN).

`torchvision.r3d_18` is taken from torchvision itself, as the reference does
(recognizers/base_moco.py:82-94).  `resnet_flow.*` is the reference's slim flow backbone
(backbones/fastonly.py:238-466): a torchvision-style VideoResNet whose width starts at 16
(basic blocks) or 8 (bottlenecks) and whose stem strides time by 2.  Written here as one
table-driven builder; parameter names match the reference's state_dict
(`stem.0.weight`, `layer1.0.conv1.0.weight`, `layer2.0.downsample.1.bias`, ...), which is an
API for downstream fine-tuning (configs/recognition/ssl_test/test_ssv2_r18.py:24-27).
"""
import torch
import torch.nn as nn


def _conv(kind, cin, cout, stride=1):
    """kind: '3d' (3x3x3, strides time), 'nt' (1x3x3, no temporal), 'nd' (3x3x3, time stride 1)."""
    if kind == "3d":
        return nn.Conv3d(cin, cout, (3, 3, 3), stride=stride, padding=1, bias=False)
    if kind == "nt":
        return nn.Conv3d(cin, cout, (1, 3, 3), stride=(1, stride, stride), padding=(0, 1, 1), bias=False)
    if kind == "nd":
        return nn.Conv3d(cin, cout, (3, 3, 3), stride=(1, stride, stride), padding=1, bias=False)
    raise ValueError(kind)


def _ds_stride(kind, stride):
    return (stride, stride, stride) if kind == "3d" else (1, stride, stride)


class _Basic(nn.Module):
    expansion = 1

    def __init__(self, cin, planes, kind, stride=1, downsample=None):
        super().__init__()
        self.conv1 = nn.Sequential(_conv(kind, cin, planes, stride), nn.BatchNorm3d(planes), nn.ReLU(inplace=True))
        self.conv2 = nn.Sequential(_conv(kind, planes, planes), nn.BatchNorm3d(planes))
        self.relu = nn.ReLU(inplace=True)
        self.downsample = downsample

    def forward(self, x):
        res = x if self.downsample is None else self.downsample(x)
        return self.relu(self.conv2(self.conv1(x)) + res)


class _Bottleneck(nn.Module):
    expansion = 4

    def __init__(self, cin, planes, kind, stride=1, downsample=None):
        super().__init__()
        self.conv1 = nn.Sequential(nn.Conv3d(cin, planes, 1, bias=False), nn.BatchNorm3d(planes), nn.ReLU(inplace=True))
        self.conv2 = nn.Sequential(_conv(kind, planes, planes, stride), nn.BatchNorm3d(planes), nn.ReLU(inplace=True))
        self.conv3 = nn.Sequential(nn.Conv3d(planes, planes * 4, 1, bias=False), nn.BatchNorm3d(planes * 4))
        self.relu = nn.ReLU(inplace=True)
        self.downsample = downsample

    def forward(self, x):
        res = x if self.downsample is None else self.downsample(x)
        return self.relu(self.conv3(self.conv2(self.conv1(x))) + res)


class _Stem(nn.Sequential):
    def __init__(self, cin, cout, t_stride, pool=False, frame_pairs=False, skip_odd=False):
        mods = [nn.Conv3d(cin, cout, (1, 7, 7), stride=(t_stride, 2, 2), padding=(0, 3, 3), bias=False),
                nn.BatchNorm3d(cout), nn.ReLU(inplace=True)]
        if pool:
            mods.append(nn.MaxPool3d((1, 3, 3), stride=(1, 2, 2), padding=(0, 1, 1)))
        super().__init__(*mods)
        self.frame_pairs, self.skip_odd = frame_pairs, skip_odd

    def forward(self, x):
        if self.frame_pairs:      # r2dv2: consecutive frame pairs folded into channels (fastonly.py:207-211)
            x = x.unflatten(2, (x.shape[2] // 2, 2)).transpose(2, 3).flatten(1, 2)
        if self.skip_odd:         # r2dv3 (fastonly.py:223-225)
            x = x[:, :, ::2]
        return super().forward(x)


# name -> (block, conv kind per stage, layers, stem kwargs)
_ARCHS = {
    "r2d_18": (_Basic, ["nt"] * 4, [2, 2, 2, 2], dict(cin=3, t_stride=2)),
    "r2dv2_18": (_Basic, ["nt"] * 4, [2, 2, 2, 2], dict(cin=6, t_stride=1, frame_pairs=True)),
    "r2dv3_18": (_Basic, ["nt"] * 4, [2, 2, 2, 2], dict(cin=3, t_stride=1, skip_odd=True)),
    "mx2d_18": (_Basic, ["nt", "nt", "nt", "3d"], [2, 2, 2, 2], dict(cin=3, t_stride=2)),
    "mc3_18": (_Basic, ["3d", "nt", "nt", "nt"], [2, 2, 2, 2], dict(cin=3, t_stride=2)),
    "r3d_18": (_Basic, ["3d"] * 4, [2, 2, 2, 2], dict(cin=3, t_stride=2)),
    "r3dv2_18": (_Basic, ["nd"] * 4, [2, 2, 2, 2], dict(cin=3, t_stride=2)),
    "r2d_50": (_Bottleneck, ["nt"] * 4, [3, 4, 6, 3], dict(cin=3, t_stride=2, pool=True)),
}


class VideoResNetSlim(nn.Module):
    """Slim VideoResNet; forward returns the four stage outputs (recognizers/moco.py:12-24)."""

    def __init__(self, name, num_classes=400):
        super().__init__()
        block, kinds, layers, stem_kw = _ARCHS[name]
        width = 16 if block is _Basic else 8
        self.stem = _Stem(cout=width, **stem_kw)
        self._cin = width
        self.layer1 = self._stage(block, kinds[0], width, layers[0], 1)
        self.layer2 = self._stage(block, kinds[1], width * 2, layers[1], 2)
        self.layer3 = self._stage(block, kinds[2], width * 4, layers[2], 2)
        self.layer4 = self._stage(block, kinds[3], width * 8, layers[3], 2)
        self.avgpool = nn.AdaptiveAvgPool3d((1, 1, 1))
        self.fc = nn.Linear(8 * width * block.expansion, num_classes)
        for m in self.modules():
            if isinstance(m, nn.Conv3d):
                nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu")
            elif isinstance(m, nn.BatchNorm3d):
                nn.init.constant_(m.weight, 1)
                nn.init.constant_(m.bias, 0)
            elif isinstance(m, nn.Linear):
                nn.init.normal_(m.weight, 0, 0.01)
                nn.init.constant_(m.bias, 0)

    def _stage(self, block, kind, planes, n, stride):
        down = None
        if stride != 1 or self._cin != planes * block.expansion:
            down = nn.Sequential(nn.Conv3d(self._cin, planes * block.expansion, 1, stride=_ds_stride(kind, stride), bias=False),
                                 nn.BatchNorm3d(planes * block.expansion))
        blocks = [block(self._cin, planes, kind, stride, down)]
        self._cin = planes * block.expansion
        blocks += [block(self._cin, planes, kind) for _ in range(1, n)]
        return nn.Sequential(*blocks)

    def forward(self, x):
        x = self.stem(x)
        outs = []
        for layer in (self.layer1, self.layer2, self.layer3, self.layer4):
            x = layer(x)
            outs.append(x)
        return outs


def ResNetFlow(name, pretrained=False, disable_clf=False, **kwargs):
    """Same call signature as backbones/fastonly.py:444."""
    if name not in _ARCHS:
        raise NotImplementedError(name)
    assert pretrained is False, "no pretrained weights for the flow backbones"
    net = VideoResNetSlim(name, **kwargs)
    if disable_clf:
        net.classifier = nn.Identity()
        net.fc = nn.Identity()
    return net


def _multilevel_forward(self, x):
    x = self.stem(x)
    outs = []
    for layer in (self.layer1, self.layer2, self.layer3, self.layer4):
        x = layer(x)
        outs.append(x)
    return outs


def torchvision_multilevel(net):
    """Make a torchvision VideoResNet return [layer1..layer4] outputs; parameter names are
    untouched (the reference monkey-patches `.forward` the same way, moco.py:374-376).
    The patch is a BOUND METHOD, not a closure: copy.deepcopy re-binds it to the copied module (a closure over `net`
    would keep running the original's layers from inside the copy, as functools.partial(forward, encoder) in the
    reference does not)."""
    import types
    net.forward = types.MethodType(_multilevel_forward, net)
    return net
