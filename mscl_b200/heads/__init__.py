from .moco_head import MoCoHead, MSCLWithAugMxHead
from .local_cl_head import MSCLWithAugPosHeadV2

__all__ = ["MoCoHead", "MSCLWithAugMxHead", "MSCLWithAugPosHeadV2"]
