from .ssl_aug import FlowVisualizer, IdentityAug, SyncMoCoAugmentV5, make_colorwheel

__all__ = ["FlowVisualizer", "IdentityAug", "SyncMoCoAugmentV5", "make_colorwheel"]
