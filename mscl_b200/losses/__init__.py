from .cross_entropy_loss import CrossEntropyLoss_torch

__all__ = ["CrossEntropyLoss_torch"]
