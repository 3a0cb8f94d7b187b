"""MSCLWithAugPosHeadV2 -- LMCL, the frame-level RGB-vs-flow contrast
(heads/local_cl_head.py:10-81), on the fused kernels.

forward(): HW-mean of q_mlvl[0] and of cat(base flow, FRA flow) along T (kernel hw_mean),
optional 1x1 Conv1d projections (PyTorch; Identity in the r18 config), then ONE kernel does
L2-normalise, t x 2t similarities / T, cross-entropy against the diagonal, top-1/5 and the
backward to the pooled features.
"""
from collections import OrderedDict

import torch
import torch.nn as nn

from .. import functional as fx
from ..registry import HEADS, build_loss
from .moco_head import topk_hits_on_device


class FusedScores:
    """What forward() hands to loss(): the kernel's [loss, top1, top5] instead of a score matrix."""

    def __init__(self, out, n_rows, n_cols):
        self.out = out
        self.shape = (n_rows, n_cols)


class _FrameContrastHead(nn.Module):
    """Shared machinery of the frame-level RGB-vs-flow contrast heads: HW-mean pooling (kernel hw_mean),
    optional 1x1 Conv1d projections (PyTorch), then ONE kernel for L2-normalise, t x t' similarities / T,
    cross-entropy against the diagonal, top-1/5 and the backward to the pooled features."""

    def __init__(self, basename, loss_cls, loss_pos, num_classes, in_channels, t, T, aux_keys):
        super().__init__()
        self.num_classes, self.in_channels = num_classes, in_channels
        self.loss_cls = build_loss(loss_cls)
        self.loss_pos = build_loss(loss_pos)
        self.multi_class, self.label_smooth_eps = False, 0.0
        self.basename = "_" + basename if basename else basename
        self.T = T
        self.aux_keys = aux_keys
        self.register_buffer("labels", torch.arange(t).unsqueeze(0))

    def init_weights(self):
        pass

    def _scores(self, x_q, x_f):
        """(b,c,t,h,w) RGB map and (b,c',t',h',w') flow map -> (FusedScores, labels)."""
        x_q = self.trans_rgb(fx.hw_mean(x_q))           # (b, 128, t); row-major and channels-last maps are read in place
        x_f = self.trans_flow(fx.hw_mean(x_f))          # (b, 128, t')
        out = fx.lmcl(x_q.contiguous(), x_f.contiguous(), self.T)
        pos_labels = self.labels.repeat((x_q.shape[0], 1)).flatten(0, 1)
        return FusedScores(out, x_q.shape[0] * x_q.shape[2], x_f.shape[2]), pos_labels

    def _loss_pos(self, pos_scores, pos_labels, scale=1.0, **kwargs):
        losses = OrderedDict()
        if isinstance(pos_scores, FusedScores):
            losses["loss_pos"] = pos_scores.out[0] * (self.loss_pos.loss_weight * scale)
            losses["top1_acc_pos"] = pos_scores.out[1].detach()
            losses["top5_acc_pos"] = pos_scores.out[2].detach()
        else:   # materialised scores (reference signature)
            losses["loss_pos"] = self.loss_pos(pos_scores, pos_labels) * scale
            top1, top5 = topk_hits_on_device(pos_scores.detach(), pos_labels)
            losses["top1_acc_pos"], losses["top5_acc_pos"] = top1, top5
        return losses

    def loss(self, pos_scores, pos_labels, **kwargs):
        return self._loss_pos(pos_scores, pos_labels)

    def update_aux_info(self, info_name, info_dict, target):
        if info_name in self.aux_keys:
            for k in self.aux_keys[info_name]:
                assert self.aux_keys[info_name][k] not in target, \
                    f"Find key-{self.aux_keys[info_name][k]} in target dict with keys:{target.keys()}"
                target[self.aux_keys[info_name][k]] = info_dict[k]
        return target


def _rgb_mlp(cin):
    return nn.Sequential(nn.Conv1d(cin, 128, 1), nn.ReLU(), nn.Conv1d(128, 128, 1))


@HEADS.register_module()
class MSCLWithAugPosHeadV2(_FrameContrastHead):
    """LMCL head of the MSCL configs (heads/local_cl_head.py:10-81): both projections optional."""

    def __init__(self, basename="", loss_cls=dict(type="CrossEntropyLoss"), loss_pos=dict(type="CrossEntropyLoss"),
                 num_classes=2, in_channels=128, mlvl_ids=(0, -1), bkb_channels=(512, 128), t=8, T=0.07,
                 aux_keys=dict()):
        super().__init__(basename, loss_cls, loss_pos, num_classes, in_channels, t, T, aux_keys)
        self.mlvl_ids = mlvl_ids
        self.trans_rgb = _rgb_mlp(bkb_channels[0]) if bkb_channels[0] is not None else nn.Identity()
        self.trans_flow = nn.Conv1d(bkb_channels[1], 128, 1) if bkb_channels[1] is not None else nn.Identity()

    def forward(self, q_mlvl, q_flow_mlvl, q_aug_flow_mlvl, **kwargs):
        x_f = torch.cat((q_flow_mlvl[self.mlvl_ids[1]], q_aug_flow_mlvl[self.mlvl_ids[1]]), dim=2)
        scores, labels = self._scores(q_mlvl[self.mlvl_ids[0]], x_f)
        return dict(pos_scores=scores, pos_labels=labels)


@HEADS.register_module()
class MSCLWithAugPosHead(_FrameContrastHead):
    """First version of the LMCL head (heads/moco_head_v2.py:197-264): the flow projection always exists."""

    def __init__(self, basename="", loss_cls=dict(type="CrossEntropyLoss"), loss_pos=dict(type="CrossEntropyLoss"),
                 num_classes=2, in_channels=128, mlvl_ids=(0, -1), bkb_channels=(512, 128), t=8, T=0.07,
                 aux_keys=dict()):
        super().__init__(basename, loss_cls, loss_pos, num_classes, in_channels, t, T, aux_keys)
        self.mlvl_ids = mlvl_ids
        self.trans_rgb = _rgb_mlp(bkb_channels[0]) if bkb_channels[0] is not None else nn.Identity()
        self.trans_flow = nn.Conv1d(bkb_channels[1], 128, 1)

    def forward(self, q_mlvl, q_flow_mlvl, q_aug_flow_mlvl, **kwargs):
        x_f = torch.cat((q_flow_mlvl[self.mlvl_ids[1]], q_aug_flow_mlvl[self.mlvl_ids[1]]), dim=2)
        scores, labels = self._scores(q_mlvl[self.mlvl_ids[0]], x_f)
        return dict(pos_scores=scores, pos_labels=labels)


@HEADS.register_module()
class MoDistv2PosHead(_FrameContrastHead):
    """Frame-level contrast without rotated-flow negatives (heads/moco_head_v2.py:128-194): t x t scores."""

    def __init__(self, basename="", loss_cls=dict(type="CrossEntropyLoss"), loss_pos=dict(type="CrossEntropyLoss"),
                 num_classes=2, in_channels=128, mlvl_ids=(0, -1), bkb_channels=(512, 128), t=8, T=0.07,
                 aux_keys=dict()):
        super().__init__(basename, loss_cls, loss_pos, num_classes, in_channels, t, T, aux_keys)
        self.mlvl_ids = mlvl_ids
        self.trans_rgb = _rgb_mlp(bkb_channels[0]) if bkb_channels[0] is not None else nn.Identity()
        self.trans_flow = nn.Conv1d(bkb_channels[1], 128, 1)

    def forward(self, q_mlvl, q_flow_mlvl, **kwargs):
        scores, labels = self._scores(q_mlvl[self.mlvl_ids[0]], q_flow_mlvl[self.mlvl_ids[1]])
        return dict(pos_scores=scores, pos_labels=labels)


@HEADS.register_module()
class MlvlMSCLWithAugPosHead(_FrameContrastHead):
    """LMCL on several pyramid levels with shared projections (heads/moco_head_v2.py:355-441): one fused
    kernel call per (rgb level, flow level) pair, each loss divided by the number of pairs, keys suffixed `_i`."""

    def __init__(self, basename="", loss_cls=dict(type="CrossEntropyLoss"), loss_pos=dict(type="CrossEntropyLoss"),
                 num_classes=2, in_channels=128, mlvl_ids=(0, 1, 2), mlvl_flow_ids=(-1, -1, -1),
                 bkb_channels=(None, None), t=8, T=0.07, pool_type="avg", aux_keys=dict()):
        super().__init__(basename, loss_cls, loss_pos, num_classes, in_channels, t, T, aux_keys)
        if pool_type != "avg":
            raise NotImplementedError(f"pool_type={pool_type!r}: only the HW-mean pooling kernel exists")
        self.mlvl_ids, self.mlvl_flow_ids = mlvl_ids, mlvl_flow_ids
        self.num_ids = len(mlvl_ids)
        self.trans_rgb = nn.Conv1d(bkb_channels[0], 128, 1) if bkb_channels[0] is not None else nn.Identity()
        self.trans_flow = nn.Conv1d(bkb_channels[1], 128, 1) if bkb_channels[1] is not None else nn.Identity()

    def forward(self, q_mlvl, q_flow_mlvl, q_aug_flow_mlvl=None, **kwargs):
        pos_scores, pos_labels = [], []
        for rgb_id, flow_id in zip(self.mlvl_ids, self.mlvl_flow_ids):
            x_f = q_flow_mlvl[flow_id]
            if q_aug_flow_mlvl is not None:
                x_f = torch.cat((x_f, q_aug_flow_mlvl[flow_id]), dim=2)
            scores, labels = self._scores(q_mlvl[rgb_id], x_f)
            pos_scores.append(scores)
            pos_labels.append(labels)
        return dict(pos_scores=pos_scores, pos_labels=pos_labels)

    def loss(self, pos_scores, pos_labels, **kwargs):
        losses = OrderedDict()
        for i, (scores, labels) in enumerate(zip(pos_scores, pos_labels)):
            for k, v in self._loss_pos(scores, labels, scale=1.0 / self.num_ids).items():
                losses[f"{k}_{i}"] = v
        return losses


@HEADS.register_module()
class MSCLWithAugSimpleHead(nn.Module):
    """Placeholder head that contributes nothing (heads/moco_head_v2.py:109-125)."""

    def __init__(self, loss_cls=dict(type="CrossEntropyLoss"), num_classes=2, in_channels=128):
        super().__init__()
        self.num_classes, self.in_channels = num_classes, in_channels
        self.loss_cls = build_loss(loss_cls)

    def init_weights(self):
        pass

    def loss(self, **kwargs):
        return dict()

    def forward(self, **kwargs):
        return dict()

    def update_aux_info(self, info_name, info_dict, target):
        return target
