from .video_resnet import ResNetFlow, VideoResNetSlim, torchvision_multilevel
from .slowonly import ResNet3dSlowOnly

__all__ = ["ResNetFlow", "VideoResNetSlim", "torchvision_multilevel", "ResNet3dSlowOnly"]
