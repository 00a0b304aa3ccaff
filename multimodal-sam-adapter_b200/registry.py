"""Minimal stand-ins for the mmseg 0.20.2 registries and mmcv Config the reference is driven by
(`mmseg.models.builder.BACKBONES/HEADS/SEGMENTORS`, `mmcv.Config.fromfile`) so that the reference's
config files build this package's modules unchanged. If the real mmseg is importable the classes are
ALSO registered there (see register_with_mmseg)."""
import os
import types


class Registry:
    def __init__(self, name):
        self.name = name
        self.module_dict = {}

    def register_module(self, name=None, force=False, module=None):
        def deco(cls):
            key = name or cls.__name__
            if key in self.module_dict and not force:
                raise KeyError(f"{key} is already registered in {self.name}")
            self.module_dict[key] = cls
            return cls
        return deco(module) if module is not None else deco

    def get(self, key):
        return self.module_dict.get(key)

    def build(self, cfg, **default_args):
        cfg = dict(cfg)
        cfg.pop("_delete_", None)
        typ = cfg.pop("type")
        cls = self.get(typ) if isinstance(typ, str) else typ
        if cls is None:
            raise KeyError(f"{typ} is not in the {self.name} registry")
        for k, v in default_args.items():
            cfg.setdefault(k, v)
        return cls(**cfg)


BACKBONES = Registry("backbone")
HEADS = Registry("head")
SEGMENTORS = Registry("segmentor")


def build_backbone(cfg):
    return BACKBONES.build(cfg)


def build_head(cfg):
    return HEADS.build(cfg)


def build_segmentor(cfg, train_cfg=None, test_cfg=None):
    return SEGMENTORS.build(cfg)


def register_with_mmseg():
    """Drop-in hook: put our classes under the reference's registry names inside a real mmseg install."""
    from mmseg.models.builder import BACKBONES as B, HEADS as H, SEGMENTORS as S  # noqa
    for src, dst in ((BACKBONES, B), (HEADS, H), (SEGMENTORS, S)):
        for k, v in src.module_dict.items():
            dst.register_module(name=k, force=True, module=v)


class ConfigDict(dict):
    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v


def _to_cd(x):
    if isinstance(x, dict):
        return ConfigDict({k: _to_cd(v) for k, v in x.items()})
    if isinstance(x, (list, tuple)):
        return type(x)(_to_cd(v) for v in x)
    return x


def _merge(base, new):
    out = dict(base)
    for k, v in new.items():
        if isinstance(v, dict) and v.get("_delete_", False):
            out[k] = {kk: vv for kk, vv in v.items() if kk != "_delete_"}
        elif isinstance(v, dict) and isinstance(out.get(k), dict):
            out[k] = _merge(out[k], v)
        else:
            out[k] = v
    return out


def load_config(path):
    """mmcv.Config.fromfile semantics needed by the reference configs: python file, `_base_` list
    (merged left to right), dict-merge with `_delete_=True` replacement."""
    path = os.path.abspath(path)
    ns = {"__file__": path}
    with open(path) as f:
        exec(compile(f.read(), path, "exec"), ns)
    cfg = {k: v for k, v in ns.items() if not k.startswith("__") and not isinstance(v, types.ModuleType)
           and not callable(v)}
    bases = cfg.pop("_base_", [])
    if isinstance(bases, str):
        bases = [bases]
    merged = {}
    for b in bases:
        merged = _merge(merged, load_config(os.path.join(os.path.dirname(path), b)))
    return _to_cd(_merge(merged, cfg))
