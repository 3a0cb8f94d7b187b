"""mscl_b200 -- B200-native implementation of the MSCL contrastive hot path.

Registry names, constructor signatures, `train_step` contract and state_dict keys are those
of megvii-research/MSCL (an MMAction2 fork) so that
configs/recognition/moco/mscl_r18_cosm_lr2e-2.py builds and trains unchanged; the
contrastive objective runs in hand-written sm_100a kernels behind a C ABI
(include/mscl_b200.h, mscl_b200/csrc).  See DESIGN.md and INTEGRATION.md.
"""
from .registry import (BACKBONES, HEADS, LOSSES, MODELS, NECKS, PIPELINES, RECOGNIZERS, SSL_AUGS, build_backbone,
                       build_head, build_loss, build_model, build_neck, build_recognizer, build_ssl_aug)
from .config import Config
from . import losses, necks, backbones, common, heads, recognizers  # noqa: F401  (populate the registries)
from .recognizers import MoCo, MoCoV2, MoDist, MSCL, MSCLWithAug

__version__ = "0.1.0"
