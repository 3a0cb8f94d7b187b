from .base_moco import BaseMoCoRecognizer
from .moco import MoCo, MoCoV2, concat_all_gather
from .mscl import MSCL, MSCLWithAug
from .modist import MoDist

__all__ = ["BaseMoCoRecognizer", "MoCo", "MoCoV2", "MSCL", "MSCLWithAug", "MoDist", "concat_all_gather"]
