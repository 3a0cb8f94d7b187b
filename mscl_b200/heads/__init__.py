from .moco_head import MoCoHead, MoCoHeadV2, MSCLWithAugMxHead
from .local_cl_head import (MSCLWithAugPosHeadV2, MSCLWithAugPosHead, MoDistv2PosHead, MlvlMSCLWithAugPosHead,
                            MSCLWithAugSimpleHead)

__all__ = ["MoCoHead", "MoCoHeadV2", "MSCLWithAugMxHead", "MSCLWithAugPosHeadV2", "MSCLWithAugPosHead",
           "MoDistv2PosHead", "MlvlMSCLWithAugPosHead", "MSCLWithAugSimpleHead"]
