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


@HEADS.register_module()
class MSCLWithAugPosHeadV2(nn.Module):
    def __init__(self, basename="", loss_cls=dict(type="CrossEntropyLoss"), loss_pos=dict(type="CrossEntropyLoss"),
                 num_classes=2, in_channels=128, mlvl_ids=(0, -1), bkb_channels=(512, 128), t=8, T=0.07,
                 aux_keys=dict()):
        super().__init__()
        self.num_classes, self.in_channels = num_classes, in_channels
        self.loss_cls = build_loss(loss_cls)
        self.loss_pos = build_loss(loss_pos)
        self.multi_class, self.label_smooth_eps = False, 0.0
        self.basename = "_" + basename if basename else basename
        self.T = T
        self.aux_keys = aux_keys
        self.mlvl_ids = mlvl_ids
        if bkb_channels[0] is not None:
            self.trans_rgb = nn.Sequential(nn.Conv1d(bkb_channels[0], 128, 1), nn.ReLU(), nn.Conv1d(128, 128, 1))
        else:
            self.trans_rgb = nn.Identity()
        self.trans_flow = nn.Conv1d(bkb_channels[1], 128, 1) if bkb_channels[1] is not None else nn.Identity()
        self.register_buffer("labels", torch.arange(t).unsqueeze(0))

    def init_weights(self):
        pass

    def forward(self, q_mlvl, q_flow_mlvl, q_aug_flow_mlvl, **kwargs):
        x_q = q_mlvl[self.mlvl_ids[0]]
        x_f = torch.cat((q_flow_mlvl[self.mlvl_ids[1]], q_aug_flow_mlvl[self.mlvl_ids[1]]), dim=2)
        x_q = fx.hw_mean(x_q.contiguous())           # (b, c, t)
        x_f = fx.hw_mean(x_f.contiguous())           # (b, c, 2t)
        x_q = self.trans_rgb(x_q)
        x_f = self.trans_flow(x_f)
        out = fx.lmcl(x_q.contiguous(), x_f.contiguous(), self.T)
        pos_labels = self.labels.repeat((x_q.shape[0], 1)).flatten(0, 1)
        return dict(pos_scores=FusedScores(out, x_q.shape[0] * x_q.shape[2], x_f.shape[2]), pos_labels=pos_labels)

    def _loss_pos(self, pos_scores, pos_labels, **kwargs):
        losses = OrderedDict()
        if isinstance(pos_scores, FusedScores):
            losses["loss_pos"] = pos_scores.out[0] * self.loss_pos.loss_weight
            losses["top1_acc_pos"] = pos_scores.out[1].detach()
            losses["top5_acc_pos"] = pos_scores.out[2].detach()
        else:   # materialised scores (reference signature)
            losses["loss_pos"] = self.loss_pos(pos_scores, pos_labels)
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
