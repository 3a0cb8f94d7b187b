from .ssl_aug import (FlowVisualizer, IdentityAug, MoCoAugmentV2, SyncMoCoAugmentV2, SyncMoCoAugmentV5,
                      make_colorwheel)

__all__ = ["FlowVisualizer", "IdentityAug", "MoCoAugmentV2", "SyncMoCoAugmentV2", "SyncMoCoAugmentV5",
           "make_colorwheel"]
