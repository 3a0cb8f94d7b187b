"""CrossEntropyLoss_torch under the reference's name (losses/cross_entropy_loss.py:122-138).

On the fused path the heads never materialise logits: the InfoNCE / LMCL kernels return the
mean cross-entropy directly and only `loss_weight` / `ignore_index` are read from this
object.  `forward` keeps the reference semantics for callers that do hold logits.
"""
import torch
import torch.nn.functional as F

from ..registry import LOSSES


@LOSSES.register_module()
class CrossEntropyLoss_torch(torch.nn.modules.CrossEntropyLoss):
    def __init__(self, weight=None, size_average=None, ignore_index=-100, reduce=None, reduction="mean",
                 loss_weight=1.0):
        super().__init__(weight, size_average, ignore_index, reduce, reduction)
        self.loss_weight = loss_weight

    def fusable(self):
        """True when the fused kernels compute exactly this loss (mean CE, no class weights)."""
        return self.weight is None and self.reduction == "mean"

    def forward(self, input, target):
        assert self.weight is None or isinstance(self.weight, torch.Tensor)
        return self.loss_weight * F.cross_entropy(input, target, weight=self.weight,
                                                  ignore_index=self.ignore_index, reduction=self.reduction)


@LOSSES.register_module()
class CrossEntropyLoss(torch.nn.Module):
    """MMAction2's standard cross-entropy (losses/cross_entropy_loss.py:9-70, base.py): the default
    `loss_cls` of every head, built even where it is never evaluated (MSCLWithAugPosHeadV2)."""

    def __init__(self, loss_weight=1.0, class_weight=None):
        super().__init__()
        self.loss_weight = loss_weight
        self.class_weight = None if class_weight is None else torch.tensor(class_weight)

    def fusable(self):
        return self.class_weight is None

    def forward(self, cls_score, label, **kwargs):
        w = None if self.class_weight is None else self.class_weight.to(cls_score.device)
        if cls_score.size() == label.size():     # soft labels
            lsm = F.log_softmax(cls_score, 1)
            if w is not None:
                lsm = lsm * w.unsqueeze(0)
            loss = -(label * lsm).sum(1)
            loss = loss.sum() / (w.unsqueeze(0) * label).sum() if w is not None else loss.mean()
        else:
            loss = F.cross_entropy(cls_score, label, weight=w, **kwargs)
        return loss * self.loss_weight
