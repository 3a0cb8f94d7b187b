"""Plug-in for an environment that HAS mmcv + the reference's `mmaction` package installed:

    custom_imports = dict(imports=['mscl_b200.mmcv_plugin'], allow_failed_imports=False)

in the config (or a plain `import mscl_b200.mmcv_plugin` before `build_model`) re-registers this
repo's classes into the reference's own registries (mmaction/models/builder.py:9-16) under the
same names with force=True, so `tools/train.py` builds the B200 path from the unchanged config.
Importing this module without mmaction raises ImportError (it is never imported by the package).
"""
import mscl_b200

MODEL_NAMES = ("MSCLWithAug", "MSCL", "MoDist", "MoCoV2", "MoCo", "MoCoHead", "MoCoHeadV2", "MSCLWithAugMxHead",
               "MSCLWithAugPosHeadV2", "MSCLWithAugPosHead", "MoDistv2PosHead", "MlvlMSCLWithAugPosHead",
               "MSCLWithAugSimpleHead", "CrossEntropyLoss_torch", "TPNMoCo", "BaseMoCo", "ResNet3dSlowOnly")
AUG_NAMES = ("SyncMoCoAugmentV5", "SyncMoCoAugmentV2", "MoCoAugmentV2", "IdentityAug")


def register_into_mmaction():
    from mmaction.models.builder import MODELS, SSL_AUGS      # the reference's registries
    for name in MODEL_NAMES:
        MODELS.register_module(name=name, force=True, module=mscl_b200.MODELS.get(name))
    for name in AUG_NAMES:
        SSL_AUGS.register_module(name=name, force=True, module=mscl_b200.SSL_AUGS.get(name))
    return MODEL_NAMES + AUG_NAMES


register_into_mmaction()
