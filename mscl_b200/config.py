"""Minimal loader for the reference's python config files.

The MSCL configs use only `_base_ = [...]` inheritance, plain assignments and dict literals
(configs/recognition/moco/mscl_r18_cosm_lr2e-2.py:1-134); tools/train.py additionally uses
`merge_from_dict`, `.get`, attribute access, `pretty_text` and `dump` (tools/train.py:82-143).
This covers exactly that surface of mmcv.Config so the config runs unchanged without mmcv.
"""
import copy
import os
import pprint


class ConfigDict(dict):
    def __getattr__(self, name):
        try:
            return self[name]
        except KeyError:
            raise AttributeError(name)

    def __setattr__(self, name, value):
        self[name] = value


def _wrap(obj):
    if isinstance(obj, dict):
        return ConfigDict({k: _wrap(v) for k, v in obj.items()})
    if isinstance(obj, (list, tuple)):
        return type(obj)(_wrap(v) for v in obj)
    return obj


def _merge(base, child):
    out = copy.deepcopy(base)
    for k, v in child.items():
        if isinstance(v, dict) and isinstance(out.get(k), dict) and not v.get("_delete_", False):
            out[k] = _merge(out[k], v)
        else:
            if isinstance(v, dict):
                v = {kk: vv for kk, vv in v.items() if kk != "_delete_"}
            out[k] = copy.deepcopy(v)
    return out


def _load_py(path):
    ns = {}
    with open(path) as f:
        exec(compile(f.read(), path, "exec"), ns)
    cfg = {k: v for k, v in ns.items() if not k.startswith("__") and not callable(v) and not isinstance(v, type(os))}
    bases = cfg.pop("_base_", [])
    if isinstance(bases, str):
        bases = [bases]
    merged = {}
    for b in bases:
        merged = _merge(merged, _load_py(os.path.join(os.path.dirname(path), b)))
    return _merge(merged, cfg)


class Config:
    def __init__(self, cfg_dict=None, filename=None):
        object.__setattr__(self, "_cfg_dict", _wrap(cfg_dict or {}))
        object.__setattr__(self, "filename", filename)

    @staticmethod
    def fromfile(filename):
        return Config(_load_py(os.path.abspath(filename)), filename)

    def __getattr__(self, name):
        return getattr(self._cfg_dict, name)

    def __setattr__(self, name, value):
        self._cfg_dict[name] = _wrap(value)

    def __getitem__(self, name):
        return self._cfg_dict[name]

    def __contains__(self, name):
        return name in self._cfg_dict

    def get(self, name, default=None):
        return self._cfg_dict.get(name, default)

    def merge_from_dict(self, options):
        nested = {}
        for full, v in options.items():
            d = nested
            keys = full.split(".")
            for k in keys[:-1]:
                d = d.setdefault(k, {})
            d[keys[-1]] = v
        object.__setattr__(self, "_cfg_dict", _wrap(_merge(self._cfg_dict, nested)))

    @property
    def pretty_text(self):
        return pprint.pformat(dict(self._cfg_dict), width=100)

    def dump(self, file=None):
        text = "\n".join(f"{k} = {pprint.pformat(v, width=100)}" for k, v in self._cfg_dict.items())
        if file is None:
            return text
        with open(file, "w") as f:
            f.write(text)
