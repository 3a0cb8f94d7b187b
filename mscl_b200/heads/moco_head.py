"""Contrastive heads under the reference's names.

MoCoHead           (heads/moco_head.py:9-81)
MoCoHeadV2         (heads/moco_head_v3.py:15-85)
MSCLWithAugMxHead  (heads/moco_head_v2.py:15-106)

In the reference these receive a materialised (N, 1+K) logits matrix, copy it to the host
and argsort it twice for top-1/5 (core/evaluation/accuracy.py:130-149).  Here the recognizer
runs the fused InfoNCE kernel and hands the head one row [loss, top1, top5] per loss term
(`loss_fused`); no logits matrix exists.  `loss(cls_score, labels)` keeps the reference's
signature for callers that do hold logits: same outputs, computed on the logits' device
without a host round trip.
"""
from collections import OrderedDict

import torch
import torch.nn as nn

from ..registry import HEADS, build_loss


def topk_hits_on_device(cls_score, labels, ks=(1, 5)):
    """top-k accuracy as 0-d tensors: the label is in the top k iff fewer than k scores exceed it
    (the reference's argsort membership test, ties aside)."""
    pos = cls_score.gather(1, labels.view(-1, 1))
    above = (cls_score > pos).sum(dim=1)
    return [(above < k).float().mean() for k in ks]


class _ContrastHead(nn.Module):
    def __init__(self, basename="", loss_cls=dict(type="CrossEntropyLoss"), num_classes=2, in_channels=128):
        super().__init__()
        self.num_classes = num_classes
        self.in_channels = in_channels
        self.loss_cls = build_loss(loss_cls)
        self.multi_class = False
        self.label_smooth_eps = 0.0
        self.basename = "_" + basename if basename else basename

    def init_weights(self):
        pass

    def _loss_from_logits(self, cls_score, labels, basename):
        losses = OrderedDict()
        if labels.shape == torch.Size([]):
            labels = labels.unsqueeze(0)
        top1, top5 = topk_hits_on_device(cls_score.detach(), labels)
        losses[f"top1_acc{basename}"] = top1
        losses[f"top5_acc{basename}"] = top5
        loss_cls = self.loss_cls(cls_score, labels)
        if isinstance(loss_cls, dict):
            losses.update(loss_cls)
        else:
            losses[f"loss_cls{basename}"] = loss_cls
        return losses

    def loss_fused(self, row, basename=None):
        """row: tensor [loss, top1, top5, 0] from functional.infonce for this loss term."""
        basename = self.basename if basename is None else basename
        losses = OrderedDict()
        losses[f"top1_acc{basename}"] = row[1].detach()
        losses[f"top5_acc{basename}"] = row[2].detach()
        losses[f"loss_cls{basename}"] = row[0] * self.loss_cls.loss_weight
        return losses

    def can_fuse(self):
        return hasattr(self.loss_cls, "fusable") and self.loss_cls.fusable()


@HEADS.register_module()
class MoCoHead(_ContrastHead):
    def forward(self, **kwargs):
        return dict()

    def loss(self, cls_score, labels, basename=None, **kwargs):
        return self._loss_from_logits(cls_score, labels, self.basename if basename is None else basename)

    def loss_mx(self, **kwargs):
        return dict()


@HEADS.register_module()
class MoCoHeadV2(_ContrastHead):
    """MoCoHead that also owns the logits step and its temperature (heads/moco_head_v3.py:15-85).

    `forward(q, k, weight)` keeps the reference's materialised form ((N,1+K) logits against a decayed snapshot);
    `forward_fused(q, k, recognizer)` runs the same term through the recognizer's fused queue pass and returns the
    [loss, top1, top5] row for `loss_fused`."""

    def __init__(self, basename="", loss_cls=dict(type="CrossEntropyLoss"), num_classes=2, in_channels=128, T=0.07):
        super().__init__(basename, loss_cls, num_classes, in_channels)
        self.T = T

    def forward(self, q, k, weight, **kwargs):
        logits = torch.cat([(q * k).sum(1, keepdim=True), q @ weight], dim=1) / self.T
        ssl_label = torch.zeros(logits.shape[0], dtype=torch.long, device=logits.device)
        return dict(cls_score=logits, ssl_label=ssl_label)

    def forward_fused(self, q, k, recognizer, dup_slot=None):
        return recognizer.contrast([(q, k, dup_slot)], self.T)[0]

    def loss(self, cls_score, ssl_label, basename=None, **kwargs):
        return self._loss_from_logits(cls_score, ssl_label, self.basename if basename is None else basename)

    def loss_mx(self, **kwargs):
        return dict()


@HEADS.register_module()
class MSCLWithAugMxHead(_ContrastHead):
    def __init__(self, basename="", loss_cls=dict(type="CrossEntropyLoss"), num_classes=2, in_channels=128,
                 same_kn=True, T=0.07):
        super().__init__(basename, loss_cls, num_classes, in_channels)
        self.same_kn = same_kn
        self.T = T

    def _forward_moco_mx(self, q, k, q_flow, k_flow, weight, weight_flow):
        """Materialised form with the reference's signature (heads/moco_head_v2.py:38-53); the fused
        recognizer does not call it."""
        rf_pos = (q * k_flow).sum(1, keepdim=True)
        fr_pos = (q_flow * k).sum(1, keepdim=True)
        w_rf, w_fr = (weight_flow, weight) if self.same_kn else (weight, weight_flow)
        rf_logits = torch.cat([rf_pos, q @ w_rf], dim=1) / self.T
        fr_logits = torch.cat([fr_pos, q_flow @ w_fr], dim=1) / self.T
        ssl_label = torch.zeros(rf_logits.shape[0], dtype=torch.long, device=rf_logits.device)
        return rf_logits, fr_logits, ssl_label

    def _loss_mx(self, cls_score, labels, basename=None, **kwargs):
        return self._loss_from_logits(cls_score, labels, self.basename if basename is None else basename)

    def loss(self, rf_logits, fr_logits, ssl_label, suffix=""):
        losses = self._loss_mx(rf_logits, ssl_label, basename=self.basename + suffix)
        losses.update(self._loss_mx(fr_logits, ssl_label, basename=self.basename + "_r" + suffix))
        return losses

    def loss_fused_mx(self, rf_row, fr_row, suffix=""):
        losses = self.loss_fused(rf_row, self.basename + suffix)
        losses.update(self.loss_fused(fr_row, self.basename + "_r" + suffix))
        return losses

    def forward(self, **kwargs):
        pass
