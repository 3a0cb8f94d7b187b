"""ResNet3dSlowOnly -- the RGB backbone of the r50 configs (`mscl_r50_cosm_lr3e-2.py`,
`moco_r50_consistent_augmentation_lr3e-2.py`), as plain PyTorch modules.

Behaviour and state_dict keys follow the reference's `ResNet3dSlowOnly` -> `ResNet3dPathway(lateral=False)` ->
`ResNet3d` chain (backbones/resnet3d_slowonly.py:16-52, resnet3d_slowfast.py:39-200, resnet3d.py:162-320,380-520,
757-862) for the bottleneck depths: every conv is an mmcv `ConvModule` there, i.e. a `.conv` / `.bn` pair, so
checkpoints keyed `conv1.conv.weight`, `layer3.2.conv2.bn.running_mean`, `layer2.0.downsample.conv.weight` ... load
unchanged.  The backbone stays in PyTorch (cuDNN); only its parameter list is on the hot path (the momentum EMA of
BASELINE config 5: 31,672,128 elements in 159 tensors).
"""
import torch.nn as nn

from ..registry import BACKBONES


class _ConvBN(nn.Module):
    """conv (no bias) + BatchNorm3d (+ ReLU), attribute names as mmcv's ConvModule."""

    def __init__(self, cin, cout, kernel, stride=1, padding=0, act=True):
        super().__init__()
        self.conv = nn.Conv3d(cin, cout, kernel, stride=stride, padding=padding, bias=False)
        self.bn = nn.BatchNorm3d(cout)
        self.activate = nn.ReLU(inplace=True) if act else None

    def forward(self, x):
        x = self.bn(self.conv(x))
        return self.activate(x) if self.activate is not None else x


class _Bottleneck3d(nn.Module):
    expansion = 4

    def __init__(self, cin, planes, spatial_stride=1, temporal_stride=1, downsample=None, inflate=True,
                 inflate_style="3x1x1"):
        super().__init__()
        # style='pytorch': the strides sit on the 3x3 convolution
        if inflate and inflate_style == "3x1x1":
            k1, p1, k2, p2 = (3, 1, 1), (1, 0, 0), (1, 3, 3), (0, 1, 1)
        elif inflate:
            k1, p1, k2, p2 = (1, 1, 1), (0, 0, 0), (3, 3, 3), (1, 1, 1)
        else:
            k1, p1, k2, p2 = (1, 1, 1), (0, 0, 0), (1, 3, 3), (0, 1, 1)
        self.conv1 = _ConvBN(cin, planes, k1, 1, p1)
        self.conv2 = _ConvBN(planes, planes, k2, (temporal_stride, spatial_stride, spatial_stride), p2)
        self.conv3 = _ConvBN(planes, planes * self.expansion, 1, act=False)
        self.downsample = downsample
        self.relu = nn.ReLU(inplace=True)

    def forward(self, x):
        identity = x if self.downsample is None else self.downsample(x)
        return self.relu(self.conv3(self.conv2(self.conv1(x))) + identity)


_STAGES = {50: (3, 4, 6, 3), 101: (3, 4, 23, 3), 152: (3, 8, 36, 3)}


@BACKBONES.register_module()
class ResNet3dSlowOnly(nn.Module):
    def __init__(self, depth, pretrained=None, stage_blocks=None, pretrained2d=True, in_channels=3, num_stages=4,
                 base_channels=64, out_indices=(3,), spatial_strides=(1, 2, 2, 2), temporal_strides=(1, 1, 1, 1),
                 conv1_kernel=(1, 7, 7), conv1_stride_s=2, conv1_stride_t=1, pool1_stride_s=2, pool1_stride_t=1,
                 with_pool1=True, with_pool2=False, inflate=(0, 0, 1, 1), inflate_style="3x1x1", lateral=False,
                 zero_init_residual=True, norm_eval=False, frozen_stages=-1, **kwargs):
        super().__init__()
        if depth not in _STAGES:
            raise NotImplementedError(f"ResNet3dSlowOnly depth {depth}: only the bottleneck depths {sorted(_STAGES)} exist here")
        if lateral:
            raise NotImplementedError("lateral connections belong to SlowFast, not to SlowOnly")
        if pretrained is not None:
            raise NotImplementedError("load pretrained weights with load_state_dict(); the keys are the reference's")
        if kwargs:
            raise TypeError(f"unsupported ResNet3dSlowOnly arguments {sorted(kwargs)}")
        assert 1 <= num_stages <= 4 and max(out_indices) < num_stages
        blocks = tuple(stage_blocks) if stage_blocks is not None else _STAGES[depth][:num_stages]
        assert len(spatial_strides) == len(temporal_strides) == num_stages == len(blocks)
        inflate = (inflate,) * num_stages if isinstance(inflate, int) else tuple(inflate)
        self.out_indices, self.with_pool1, self.with_pool2 = tuple(out_indices), with_pool1, with_pool2
        self.zero_init_residual, self.norm_eval, self.frozen_stages = zero_init_residual, norm_eval, frozen_stages
        kt = (conv1_kernel,) * 3 if isinstance(conv1_kernel, int) else tuple(conv1_kernel)
        self.conv1 = _ConvBN(in_channels, base_channels, kt, (conv1_stride_t, conv1_stride_s, conv1_stride_s),
                             tuple((k - 1) // 2 for k in kt))
        self.maxpool = nn.MaxPool3d((1, 3, 3), (pool1_stride_t, pool1_stride_s, pool1_stride_s), (0, 1, 1))
        self.pool2 = nn.MaxPool3d((2, 1, 1), (2, 1, 1))
        self.res_layers = []
        cin = base_channels
        for i, n in enumerate(blocks):
            planes = base_channels * 2 ** i
            infl = (inflate[i],) * n if isinstance(inflate[i], int) else tuple(inflate[i])
            ss, ts = spatial_strides[i], temporal_strides[i]
            down = None
            if ss != 1 or cin != planes * _Bottleneck3d.expansion:
                down = _ConvBN(cin, planes * _Bottleneck3d.expansion, 1, (ts, ss, ss), act=False)
            layers = [_Bottleneck3d(cin, planes, ss, ts, down, infl[0] == 1, inflate_style)]
            cin = planes * _Bottleneck3d.expansion
            layers += [_Bottleneck3d(cin, planes, inflate=infl[j] == 1, inflate_style=inflate_style) for j in range(1, n)]
            self.add_module(f"layer{i + 1}", nn.Sequential(*layers))
            self.res_layers.append(f"layer{i + 1}")
        self.feat_dim = cin

    def init_weights(self, pretrained=None):
        """Kaiming-normal (fan_out) convolutions, unit batch norms, zero last batch norm of every block
        (resnet3d.py:817-831)."""
        if pretrained is not None:
            raise NotImplementedError("load pretrained weights with load_state_dict()")
        for m in self.modules():
            if isinstance(m, nn.Conv3d):
                nn.init.kaiming_normal_(m.weight, a=0, mode="fan_out", nonlinearity="relu")
            elif isinstance(m, nn.BatchNorm3d):
                nn.init.constant_(m.weight, 1)
                nn.init.constant_(m.bias, 0)
        if self.zero_init_residual:
            for m in self.modules():
                if isinstance(m, _Bottleneck3d):
                    nn.init.constant_(m.conv3.bn.weight, 0)

    def _freeze_stages(self):
        if self.frozen_stages >= 0:
            self.conv1.eval()
            for p in self.conv1.parameters():
                p.requires_grad = False
        for i in range(1, self.frozen_stages + 1):
            m = getattr(self, f"layer{i}")
            m.eval()
            for p in m.parameters():
                p.requires_grad = False

    def train(self, mode=True):
        super().train(mode)
        self._freeze_stages()
        if mode and self.norm_eval:
            for m in self.modules():
                if isinstance(m, nn.BatchNorm3d):
                    m.eval()
        return self

    def forward(self, x):
        x = self.conv1(x)
        if self.with_pool1:
            x = self.maxpool(x)
        outs = []
        for i, name in enumerate(self.res_layers):
            x = getattr(self, name)(x)
            if i == 0 and self.with_pool2:
                x = self.pool2(x)
            if i in self.out_indices:
                outs.append(x)
        return outs[0] if len(outs) == 1 else tuple(outs)
