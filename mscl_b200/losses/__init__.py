from .cross_entropy_loss import CrossEntropyLoss, CrossEntropyLoss_torch

__all__ = ["CrossEntropyLoss", "CrossEntropyLoss_torch"]
