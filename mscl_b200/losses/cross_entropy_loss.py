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
