"""Drive the UNMODIFIED reference (/root/reference) on CPU in the build container.

Only used by tools/make_golden.py and tools/validate_oracle.py to pin the oracle and to write the
small golden fixtures under tests/golden/. Nothing in tests/, bench.py or the product imports
this at run time (the reference tree does not exist on the GPU box).

Recipe = SURVEY.md Appendix C: stub the un-installed third-party packages (mmcv/mmseg/timm/...),
pre-register bare namespace packages so heavyweight __init__ files never run, chdir into
segmentation/, neuter the checkpoint download, and route MSDeformAttnFunction to the reference's
own CPU ground truth ms_deform_attn_core_pytorch.
"""
import contextlib
import io
import os
import sys
import types

import torch
import torch.nn as nn

REF = os.environ.get("MMSAM_REFERENCE", "/root/reference")
SEG = os.path.join(REF, "segmentation")


class _Registry:
    def __init__(self, name):
        self.name, self.module_dict = name, {}

    def register_module(self, name=None, force=False, module=None):
        def deco(cls):
            self.module_dict[name or cls.__name__] = cls
            return cls
        if module is not None:
            return deco(module)
        return deco

    def get(self, k):
        return self.module_dict.get(k)

    def build(self, cfg, *a, **k):
        cfg = dict(cfg)
        return self.module_dict[cfg.pop("type")](**cfg)


def _mod(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def _ns(name, path):
    m = types.ModuleType(name)
    m.__path__ = [path]
    sys.modules[name] = m
    return m


_installed = False


def install():
    global _installed
    if _installed:
        return
    _installed = True
    sys.dont_write_bytecode = True
    os.chdir(SEG)
    sys.path.insert(0, SEG)

    class Dict(dict):
        __getattr__ = dict.get
        __setattr__ = dict.__setitem__
    _mod("addict", Dict=Dict)
    _mod("yapf"); _mod("yapf.yapflib"); _mod("yapf.yapflib.yapf_api", FormatCode=lambda *a, **k: (a[0], False))
    _mod("termcolor", colored=lambda s, *a, **k: s)
    _mod("matplotlib"); _mod("matplotlib.pyplot")

    class DropPath(nn.Module):
        def __init__(self, p=0.):
            super().__init__()
            self.drop_prob = p

        def forward(self, x):
            assert not self.training
            return x
    _mod("timm"); _mod("timm.models")
    _mod("timm.models.layers", DropPath=DropPath, drop_path=lambda x, *a, **k: x,
         to_2tuple=lambda x: x if isinstance(x, (tuple, list)) else (x, x), trunc_normal_=nn.init.trunc_normal_)
    regs = dict(BACKBONES=_Registry("backbone"), HEADS=_Registry("head"), SEGMENTORS=_Registry("segmentor"))
    _mod("mmseg"); _mod("mmseg.models"); _mod("mmseg.models.builder", **regs)
    import logging
    _mod("mmseg.utils", get_root_logger=lambda *a, **k: logging.getLogger("ref"))
    _mod("MultiScaleDeformableAttention")
    m = _ns("mmcv_custom", os.path.join(SEG, "mmcv_custom"))
    m.load_checkpoint = lambda *a, **k: None
    _ns("mmpretrain_custom", os.path.join(SEG, "mmpretrain_custom"))
    _ns("mmseg_custom", os.path.join(SEG, "mmseg_custom"))
    _ns("mmseg_custom.models", os.path.join(SEG, "mmseg_custom/models"))
    _ns("mmseg_custom.models.backbones", os.path.join(SEG, "mmseg_custom/models/backbones"))
    _ns("mmseg_custom.models.backbones.base", os.path.join(SEG, "mmseg_custom/models/backbones/base"))

    with contextlib.redirect_stdout(io.StringIO()):
        import mmengine_custom.runner as R
        R.CheckpointLoader.load_checkpoint = staticmethod(
            lambda *a, **k: {"state_dict": {"dummy.x": torch.zeros(1)}})
        import ops.modules.ms_deform_attn as M
        from ops.functions.ms_deform_attn_func import ms_deform_attn_core_pytorch

        class _Fn:
            @staticmethod
            def apply(value, shapes, lsi, loc, w, step):
                return ms_deform_attn_core_pytorch(value, shapes, loc, w)
        M.MSDeformAttnFunction = _Fn


def backbone_cls(withcp=False):
    install()
    with contextlib.redirect_stdout(io.StringIO()):
        if withcp:
            from mmseg_custom.models.backbones.image_encoder_adapter_bimodal_mix_mod_new_in_twin_convnext_new_with_cp import \
                SAMAdapterbimodalMixModNewInTwinConvNEWwithcp as C
        else:
            from mmseg_custom.models.backbones.image_encoder_adapter_bimodal_mix_mod_new_in_twin_convnext_new import \
                SAMAdapterbimodalMixModNewInTwinConvNEW as C
    return C


def build_backbone(cfg, withcp=False):
    C = backbone_cls(withcp)
    cfg = {k: v for k, v in cfg.items() if k not in ("type", "_delete_")}
    with contextlib.redirect_stdout(io.StringIO()):
        net = C(**cfg)
    return net.eval()


def ref_msda_core():
    install()
    from ops.functions.ms_deform_attn_func import ms_deform_attn_core_pytorch
    return ms_deform_attn_core_pytorch
