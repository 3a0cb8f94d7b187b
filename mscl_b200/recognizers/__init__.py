from .base_moco import BaseMoCoRecognizer
from .moco import MoCoV2, concat_all_gather
from .mscl import MSCLWithAug

__all__ = ["BaseMoCoRecognizer", "MoCoV2", "MSCLWithAug", "concat_all_gather"]
