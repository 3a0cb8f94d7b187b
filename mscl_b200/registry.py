"""Registries and builders with the reference's names.

Mirror of `mmaction/models/builder.py:9-97` (MODELS and its aliases, SSL_AUGS, build_*)
for an environment without mmcv: `type=` strings of the MSCL configs resolve here.  When
mmcv/mmaction ARE importable, `import mscl_b200.mmcv_plugin` registers
the same classes into the reference's own registry instead.
"""
import warnings


class Registry:
    """Name -> class map; supports `@R.register_module()`, bare `@R.register_module`
    (used by the reference at necks/sepc.py:16) and `R.build(cfg, default_args)`."""

    def __init__(self, name, parent=None):
        self.name = name
        self.parent = parent
        self.module_dict = {}

    def __len__(self):
        return len(self.module_dict)

    def __contains__(self, key):
        return self.get(key) is not None

    def get(self, key):
        if key in self.module_dict:
            return self.module_dict[key]
        return self.parent.get(key) if self.parent is not None else None

    def _register(self, cls, name=None, force=False):
        key = name or cls.__name__
        if key in self.module_dict and not force:
            raise KeyError(f"{key} is already registered in {self.name}")
        self.module_dict[key] = cls
        return cls

    def register_module(self, name=None, force=False, module=None):
        if isinstance(name, type):      # bare decorator
            return self._register(name)
        if module is not None:
            return self._register(module, name, force)
        return lambda cls: self._register(cls, name, force)

    def build(self, cfg, default_args=None):
        if not isinstance(cfg, dict) or "type" not in cfg:
            raise TypeError(f"cfg must be a dict with a 'type' key, got {cfg!r}")
        args = dict(cfg)
        for k, v in (default_args or {}).items():
            args.setdefault(k, v)
        typ = args.pop("type")
        cls = self.get(typ) if isinstance(typ, str) else typ
        if cls is None:
            raise KeyError(f"{typ} is not in the {self.name} registry")
        return cls(**args)


MODELS = Registry("models")
BACKBONES = NECKS = HEADS = RECOGNIZERS = LOSSES = LOCALIZERS = MODELS
SSL_AUGS = Registry("ssl_augs")
PIPELINES = Registry("pipelines")


def build_backbone(cfg):
    return BACKBONES.build(cfg)


def build_head(cfg):
    return HEADS.build(cfg)


def build_neck(cfg):
    return NECKS.build(cfg)


def build_loss(cfg):
    return LOSSES.build(cfg)


def build_ssl_aug(cfg):
    return SSL_AUGS.build(cfg)


def build_recognizer(cfg, train_cfg=None, test_cfg=None):
    if train_cfg is not None or test_cfg is not None:
        warnings.warn("train_cfg and test_cfg is deprecated, please specify them in model", UserWarning)
    assert cfg.get("train_cfg") is None or train_cfg is None, "train_cfg specified in both outer field and model field"
    assert cfg.get("test_cfg") is None or test_cfg is None, "test_cfg specified in both outer field and model field"
    return RECOGNIZERS.build(cfg, default_args=dict(train_cfg=train_cfg, test_cfg=test_cfg))


def build_model(cfg, train_cfg=None, test_cfg=None):
    typ = cfg["type"]
    if typ not in RECOGNIZERS:
        raise ValueError(f"{typ} is not registered in LOCALIZERS, RECOGNIZERS or DETECTORS")
    return build_recognizer(cfg, train_cfg, test_cfg)
