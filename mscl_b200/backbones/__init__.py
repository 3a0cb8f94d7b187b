from .video_resnet import ResNetFlow, VideoResNetSlim, torchvision_multilevel

__all__ = ["ResNetFlow", "VideoResNetSlim", "torchvision_multilevel"]
