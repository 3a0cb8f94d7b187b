"""mscl_b200/mmcv_plugin.py: importing it re-registers this repo's classes into the
reference's own registries (mmaction/models/builder.py:9-16) under the same names with force=True.

mmcv and the reference package are absent from this image, so the test installs a stand-in `mmaction.models.builder`
whose Registry follows mmcv 1.3's `register_module(name=None, force=False, module=None)` contract (KeyError on a
duplicate name unless force), pre-populated with placeholder classes under every name the reference registers -- exactly
the situation of a real install, where the reference's own classes are already there when the plug-in is imported."""
import importlib
import os
import sys
import types

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class _Registry:
    """mmcv.utils.Registry (1.3.x) as far as the plug-in and `build_model` use it."""

    def __init__(self, name):
        self.name, self.module_dict = name, {}

    def get(self, key):
        return self.module_dict.get(key)

    def _register_module(self, module_class, module_name=None, force=False):
        if not isinstance(module_class, type):
            raise TypeError(f"module must be a class, but got {type(module_class)}")
        name = module_name or module_class.__name__
        if not force and name in self.module_dict:
            raise KeyError(f"{name} is already registered in {self.name}")
        self.module_dict[name] = module_class

    def register_module(self, name=None, force=False, module=None):
        if module is not None:
            self._register_module(module, name, force)
            return module

        def deco(cls):
            self._register_module(cls, name, force)
            return cls
        return deco

    def build(self, cfg, default_args=None):
        args = dict(cfg)
        for k, v in (default_args or {}).items():
            args.setdefault(k, v)
        return self.module_dict[args.pop("type")](**args)


@pytest.fixture
def fake_mmaction():
    saved = {k: sys.modules.get(k) for k in ("mmaction", "mmaction.models", "mmaction.models.builder", "mscl_b200.mmcv_plugin")}
    builder = types.ModuleType("mmaction.models.builder")
    builder.MODELS = _Registry("models")
    builder.SSL_AUGS = _Registry("ssl_augs")
    builder.RECOGNIZERS = builder.HEADS = builder.NECKS = builder.LOSSES = builder.BACKBONES = builder.MODELS
    mm, models = types.ModuleType("mmaction"), types.ModuleType("mmaction.models")
    mm.models, models.builder = models, builder
    sys.modules.update({"mmaction": mm, "mmaction.models": models, "mmaction.models.builder": builder})
    sys.modules.pop("mscl_b200.mmcv_plugin", None)
    yield builder
    for k, v in saved.items():
        if v is None:
            sys.modules.pop(k, None)
        else:
            sys.modules[k] = v


def test_plugin_overrides_the_reference_registrations(fake_mmaction):
    import mscl_b200
    builder = fake_mmaction
    # the reference's own classes are registered first (import order of a real install)
    names_models = ("MSCLWithAug", "MSCL", "MoDist", "MoCoV2", "MoCo", "MoCoHead", "MoCoHeadV2", "MSCLWithAugMxHead",
                    "MSCLWithAugPosHeadV2", "MSCLWithAugPosHead", "MoDistv2PosHead", "MlvlMSCLWithAugPosHead",
                    "MSCLWithAugSimpleHead", "CrossEntropyLoss_torch", "TPNMoCo", "BaseMoCo", "ResNet3dSlowOnly")
    names_augs = ("SyncMoCoAugmentV5", "SyncMoCoAugmentV2", "MoCoAugmentV2", "IdentityAug")
    placeholders = {}
    for n in names_models:
        placeholders[n] = builder.MODELS.register_module(name=n, module=type(n, (), {"__module__": "mmaction.reference"}))
    for n in names_augs:
        placeholders[n] = builder.SSL_AUGS.register_module(name=n, module=type(n, (), {"__module__": "mmaction.reference"}))
    with pytest.raises(KeyError):       # the stand-in registry refuses duplicates like mmcv's does
        builder.MODELS.register_module(name="MoCoV2", module=placeholders["MoCoV2"])
    plugin = importlib.import_module("mscl_b200.mmcv_plugin")
    assert set(plugin.MODEL_NAMES) == set(names_models) and set(plugin.AUG_NAMES) == set(names_augs)
    for n in names_models:
        cls = builder.MODELS.get(n)
        assert cls is mscl_b200.MODELS.get(n) and cls is not placeholders[n] and cls.__module__.startswith("mscl_b200."), n
    for n in names_augs:
        cls = builder.SSL_AUGS.get(n)
        assert cls is mscl_b200.SSL_AUGS.get(n) and cls.__module__.startswith("mscl_b200."), n


def test_reference_config_builds_the_b200_classes_through_the_reference_registry(fake_mmaction):
    """`tools/train.py`'s `build_model(cfg.model)` on the UNCHANGED r18 config model dict, resolved through the
    reference-side registry after the plug-in import, constructs this repo's recognizer with its heads and aug."""
    import mscl_b200
    from mscl_b200.configs import mscl_r18_model
    builder = fake_mmaction
    importlib.import_module("mscl_b200.mmcv_plugin")
    cfg = mscl_r18_model(K=256)
    model = builder.MODELS.build(cfg)
    assert type(model) is mscl_b200.MODELS.get("MSCLWithAug")
    assert type(model.recognizer) is mscl_b200.MODELS.get("MoCoV2")
    assert type(model.moco_mx_head) is mscl_b200.MODELS.get("MSCLWithAugMxHead")
    assert type(model.sup_head) is mscl_b200.MODELS.get("MSCLWithAugPosHeadV2")
    assert type(model.aug_gpu) is mscl_b200.SSL_AUGS.get("SyncMoCoAugmentV5")
    keys = model.state_dict().keys()
    assert {"recognizer.queue", "recognizer.queue_ptr", "recognizer.count", "recognizer_flow.queue"} <= set(keys)


def test_plugin_needs_the_reference_package():
    """Without mmaction the import fails loudly (the package itself never imports the plug-in)."""
    if "mmaction" in sys.modules:
        pytest.skip("an mmaction module is installed here")
    sys.modules.pop("mscl_b200.mmcv_plugin", None)
    with pytest.raises(ImportError):
        importlib.import_module("mscl_b200.mmcv_plugin")
