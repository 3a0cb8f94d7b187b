"""Import shim that executes the UNMODIFIED reference hot-path files from /root/reference.

TEST INFRASTRUCTURE ONLY.  Used by oracle/make_golden.py (to generate tests/golden/*)
and by tests/test_oracle_vs_reference.py (live pin, skipped when /root/reference is
absent, e.g. on the GPU box).  Nothing under mscl_b200/ imports this.

The reference tree is not importable as a package (mmcv / kornia / nori2 absent,
broken imports -- SURVEY.md section 0), but the hot-path files are pure PyTorch.
Recipe:
  * fake parent packages `mmaction`, `mmaction.models.*`, `mmaction.core`, `mmaction.utils`
  * a 15-line registry standing in for mmcv.utils.Registry
  * mmcv.runner / mmcv.cnn stubs (auto_fp16 = identity, ConvModule = conv[+bn][+act])
  * Tensor.cuda = identity on CPU (moco.py:160,484,498 hard-code .cuda())
No reference source is copied: files are read and exec'd from where they lie.
"""
import importlib.util
import os
import sys
import types

import torch
import torch.nn as nn

REF_ROOT = os.environ.get("MSCL_REFERENCE_ROOT", "/root/reference")


def available():
    return os.path.isdir(os.path.join(REF_ROOT, "mmaction", "models", "recognizers"))


class _Registry:
    def __init__(self, name, parent=None):
        self.name = name
        self.module_dict = {}

    def __contains__(self, key):
        return key in self.module_dict

    def get(self, key):
        return self.module_dict.get(key)

    def register_module(self, cls=None, name=None, force=False, module=None):
        if cls is not None and isinstance(cls, type):  # bare @R.register_module
            self.module_dict[cls.__name__] = cls
            return cls

        def deco(c):
            self.module_dict[name or c.__name__] = c
            return c

        return deco

    def build(self, cfg, default_args=None):
        args = dict(cfg)
        if default_args:
            for k, v in default_args.items():
                args.setdefault(k, v)
        typ = args.pop("type")
        if typ not in self.module_dict:
            raise KeyError(f"{typ} is not in the {self.name} registry")
        return self.module_dict[typ](**args)


class _ConvModule(nn.Module):
    """mmcv.cnn.ConvModule subset used by the necks: conv (+norm) (+act), bias='auto'."""

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1,
                 groups=1, bias="auto", conv_cfg=None, norm_cfg=None, act_cfg=dict(type="ReLU"),
                 inplace=True, **kw):
        super().__init__()
        with_norm = norm_cfg is not None
        if bias == "auto":
            bias = not with_norm
        conv_type = (conv_cfg or {}).get("type", "Conv2d")
        conv_cls = {"Conv2d": nn.Conv2d, "Conv3d": nn.Conv3d, "Conv1d": nn.Conv1d}[conv_type]
        self.conv = conv_cls(in_channels, out_channels, kernel_size, stride=stride, padding=padding,
                             dilation=dilation, groups=groups, bias=bias)
        self.with_norm = with_norm
        if with_norm:
            bn = {"BN3d": nn.BatchNorm3d, "BN": nn.BatchNorm2d, "BN2d": nn.BatchNorm2d}[norm_cfg["type"]]
            self.bn = bn(out_channels)
        self.activate = nn.ReLU(inplace=inplace) if act_cfg is not None else None
        nn.init.kaiming_normal_(self.conv.weight, a=0, nonlinearity="relu")
        if self.conv.bias is not None:
            nn.init.constant_(self.conv.bias, 0)

    def forward(self, x):
        x = self.conv(x)
        if self.with_norm:
            x = self.bn(x)
        if self.activate is not None:
            x = self.activate(x)
        return x


def _xavier_init(module, gain=1, bias=0, distribution="normal"):
    if distribution == "uniform":
        nn.init.xavier_uniform_(module.weight, gain=gain)
    else:
        nn.init.xavier_normal_(module.weight, gain=gain)
    if getattr(module, "bias", None) is not None:
        nn.init.constant_(module.bias, bias)


def _constant_init(module, val, bias=0):
    nn.init.constant_(module.weight, val)
    if getattr(module, "bias", None) is not None:
        nn.init.constant_(module.bias, bias)


def _normal_init(module, mean=0, std=1, bias=0):
    nn.init.normal_(module.weight, mean, std)
    if getattr(module, "bias", None) is not None:
        nn.init.constant_(module.bias, bias)


def _kaiming_init(module, a=0, mode="fan_out", nonlinearity="relu", bias=0, distribution="normal"):
    if distribution == "uniform":
        nn.init.kaiming_uniform_(module.weight, a=a, mode=mode, nonlinearity=nonlinearity)
    else:
        nn.init.kaiming_normal_(module.weight, a=a, mode=mode, nonlinearity=nonlinearity)
    if getattr(module, "bias", None) is not None:
        nn.init.constant_(module.bias, bias)


def _auto_fp16(*a, **k):
    def deco(f):
        return f

    return deco


def _pkg(name):
    m = types.ModuleType(name)
    m.__path__ = []
    sys.modules[name] = m
    return m


def _load(modname, relpath, strip_lines=(), inject=None):
    path = os.path.join(REF_ROOT, relpath)
    if not strip_lines and inject is None:
        spec = importlib.util.spec_from_file_location(modname, path)
        mod = importlib.util.module_from_spec(spec)
        sys.modules[modname] = mod
        spec.loader.exec_module(mod)
        return mod
    with open(path) as f:
        lines = f.read().split("\n")
    for ln in strip_lines:  # 1-based line numbers of dead imports
        lines[ln - 1] = "pass"
    mod = types.ModuleType(modname)
    mod.__file__ = path
    mod.__package__ = modname.rpartition(".")[0]
    if inject:
        mod.__dict__.update(inject)
    sys.modules[modname] = mod
    exec(compile("\n".join(lines), path, "exec"), mod.__dict__)
    return mod


def load_defs(relpath, names, extra_globals=None):
    """Execute ONLY the named top-level functions / classes of a reference file (read from REF_ROOT at run time,
    never copied): for files whose module-level imports cannot be satisfied here (kornia in common/ssl_aug.py)."""
    import ast
    path = os.path.join(REF_ROOT, relpath)
    with open(path) as f:
        tree = ast.parse(f.read(), filename=path)
    keep = [n for n in tree.body if isinstance(n, (ast.FunctionDef, ast.ClassDef)) and n.name in names]
    missing = set(names) - {n.name for n in keep}
    if missing:
        raise RuntimeError(f"{relpath}: no top-level definition of {sorted(missing)}")
    ns = dict(extra_globals or {})
    exec(compile(ast.Module(body=keep, type_ignores=[]), path, "exec"), ns)
    return types.SimpleNamespace(**{n: ns[n] for n in names})


_LOADED = None


def load_reference():
    """Return a namespace with the reference's own classes (executed from REF_ROOT)."""
    global _LOADED
    if _LOADED is not None:
        return _LOADED
    if not available():
        raise RuntimeError(f"reference tree not found at {REF_ROOT}")
    saved = {k: sys.modules.get(k) for k in list(sys.modules) if k.split(".")[0] in ("mmaction", "mmcv")}
    if not torch.cuda.is_available():
        torch.Tensor.cuda = lambda self, *a, **k: self

    # --- mmcv stubs ---
    mmcv = _pkg("mmcv")
    runner = _pkg("mmcv.runner")
    runner.auto_fp16 = _auto_fp16
    runner._load_checkpoint = lambda *a, **k: {}
    runner.load_state_dict = lambda *a, **k: None
    cnn = _pkg("mmcv.cnn")
    cnn.ConvModule = _ConvModule
    cnn.xavier_init = _xavier_init
    cnn.constant_init = _constant_init
    cnn.normal_init = _normal_init
    cnn.MODELS = _Registry("mmcv_models")
    cnn.NonLocal3d = None
    cnn.build_activation_layer = lambda cfg: nn.ReLU(inplace=bool(cfg.get("inplace", False)))
    cnn.kaiming_init = _kaiming_init
    runner.load_checkpoint = lambda *a, **k: {}
    utils = _pkg("mmcv.utils")
    utils.Registry = _Registry
    utils._BatchNorm = nn.modules.batchnorm._BatchNorm
    utils.print_log = lambda *a, **k: None
    mmcv.runner, mmcv.cnn, mmcv.utils = runner, cnn, utils

    # --- fake mmaction parents ---
    _pkg("mmaction")
    mm_utils = _pkg("mmaction.utils")
    import logging
    mm_utils.get_root_logger = lambda *a, **k: logging.getLogger("mmaction")
    mm_utils.import_module_error_func = lambda name: (lambda f: f)
    core = _pkg("mmaction.core")
    acc = _load("mmaction.core.evaluation_accuracy", "mmaction/core/evaluation/accuracy.py")
    core.top_k_accuracy = acc.top_k_accuracy
    core.bbox_overlaps = None
    models = _pkg("mmaction.models")
    for sub in ("heads", "losses", "necks", "recognizers", "backbones"):
        _pkg(f"mmaction.models.{sub}")
    builder = _load("mmaction.models.builder", "mmaction/models/builder.py")
    models.builder = builder

    # losses
    _load("mmaction.models.losses.base", "mmaction/models/losses/base.py")
    ce = _load("mmaction.models.losses.cross_entropy_loss", "mmaction/models/losses/cross_entropy_loss.py")
    # necks
    _load("mmaction.models.necks.fpn", "mmaction/models/necks/fpn.py")
    _load("mmaction.models.necks.sepc", "mmaction/models/necks/sepc.py")
    _load("mmaction.models.necks.fpn_video", "mmaction/models/necks/fpn_video.py")
    necks = _load("mmaction.models.necks.base", "mmaction/models/necks/base.py")
    # recognizers
    _load("mmaction.models.recognizers.base", "mmaction/models/recognizers/base.py")
    _load("mmaction.models.recognizers.base_moco", "mmaction/models/recognizers/base_moco.py")
    moco = _load("mmaction.models.recognizers.moco", "mmaction/models/recognizers/moco.py")
    mscl = _load("mmaction.models.recognizers.mscl", "mmaction/models/recognizers/mscl.py")
    mscl.forward = None  # heads/moco_head_v2.py:8 imports a name that does not exist
    # backbones: fastonly.py has two dead imports (lines 3 and 5)
    fastonly = _load("mmaction.models.backbones.fastonly", "mmaction/models/backbones/fastonly.py",
                     strip_lines=(3, 5))
    # heads
    _load("mmaction.models.heads.base", "mmaction/models/heads/base.py")
    moco_head = _load("mmaction.models.heads.moco_head", "mmaction/models/heads/moco_head.py")
    moco_head_v2 = _load("mmaction.models.heads.moco_head_v2", "mmaction/models/heads/moco_head_v2.py")
    local_cl = _load("mmaction.models.heads.local_cl_head", "mmaction/models/heads/local_cl_head.py")

    # an identity 3-argument augmentation under the config's names (kornia is absent)
    class _IdentityAug3:
        def __init__(self, **kw):
            pass

        def __call__(self, im_q, im_k=None, aux_info=None):
            if im_k is None:
                return im_q
            return im_q, im_k, aux_info

    for nm in ("IdentityAug", "SyncMoCoAugmentV5"):
        builder.SSL_AUGS.module_dict[nm] = _IdentityAug3

    # FRA (numpy): exec transforms_motion.py with its two package imports replaced
    tm_path = "mmaction/datasets/pipelines/transforms_motion.py"
    pipes = _Registry("pipelines")
    tm = _load("mmaction_ref_transforms_motion", tm_path, strip_lines=(3, 4),
               inject={"PIPELINES": pipes, "flow_viz": None})

    ns = types.SimpleNamespace(
        builder=builder, MoCoV2=moco.MoCoV2, MSCLWithAug=mscl.MSCLWithAug,
        MoCoHead=moco_head.MoCoHead, MSCLWithAugMxHead=moco_head_v2.MSCLWithAugMxHead,
        MSCLWithAugPosHeadV2=local_cl.MSCLWithAugPosHeadV2,
        CrossEntropyLoss_torch=ce.CrossEntropyLoss_torch, top_k_accuracy=acc.top_k_accuracy,
        TPNMoCo=necks.TPNMoCo, BaseMoCo=necks.BaseMoCo, ResNetFlow=fastonly.ResNetFlow,
        NormFlowWithStidedAug=tm.NormFlowWithStidedAug, norm_flow=tm.norm_flow,
        concat_all_gather=moco.concat_all_gather, moco_module=moco,
    )
    _LOADED = ns
    return ns


def load_slowonly():
    """The reference's ResNet3dSlowOnly class (backbones/resnet3d_slowonly.py:16-52; its line 2 is a stray
    `from turtle import forward`, which needs tkinter, and is skipped)."""
    load_reference()
    if "mmaction.models.backbones.resnet3d_slowonly" not in sys.modules:
        _load("mmaction.models.backbones.resnet3d", "mmaction/models/backbones/resnet3d.py")
        _load("mmaction.models.backbones.resnet3d_slowfast", "mmaction/models/backbones/resnet3d_slowfast.py")
        _load("mmaction.models.backbones.resnet3d_slowonly", "mmaction/models/backbones/resnet3d_slowonly.py",
              strip_lines=(2,))
    return sys.modules["mmaction.models.backbones.resnet3d_slowonly"].ResNet3dSlowOnly


def ensure_process_group():
    """world-size-1 gloo group so the reference's all_gather / broadcast calls run."""
    import torch.distributed as dist
    if not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29517")
        dist.init_process_group("gloo", rank=0, world_size=1)
