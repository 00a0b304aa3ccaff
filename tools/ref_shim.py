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


# ------------------------------------------------------------------------------------------------
# Head + segmentor: the reference's own SegformerHead.forward
# (decode_heads/segformer_head.py:48-66) and EncoderDecoder inference glue (segmentors/encoder_decoder.py)
# run UNMODIFIED; `mmcv.cnn.ConvModule` is pointed at the copy the reference vendors
# (mmcv_custom/cnn/bricks/conv_module.py:72-212). Their base classes live in mmsegmentation 0.20.2, which
# is neither vendored nor installed: the two stubs below restate only what the inference path touches
# (BaseDecodeHead: conv_seg / dropout / _transform_inputs / cls_seg / forward_test; BaseSegmentor: the
# with_* properties).
# ------------------------------------------------------------------------------------------------
_seg_installed = False


def install_segmentor():
    global _seg_installed
    install()
    if _seg_installed:
        return
    _seg_installed = True
    import torch.nn.functional as F
    with contextlib.redirect_stdout(io.StringIO()):
        from mmcv_custom.cnn.bricks.conv_module import ConvModule
    _mod("mmcv"); _mod("mmcv.cnn", ConvModule=ConvModule)

    class BaseDecodeHead(nn.Module):
        def __init__(self, in_channels, channels, *, num_classes, dropout_ratio=0.1, conv_cfg=None, norm_cfg=None,
                     act_cfg=dict(type="ReLU"), in_index=-1, input_transform=None, loss_decode=None, ignore_index=255,
                     sampler=None, align_corners=False, init_cfg=None):
            super().__init__()
            assert input_transform == "multiple_select" and len(in_channels) == len(in_index)
            self.in_channels, self.in_index, self.input_transform = in_channels, in_index, input_transform
            self.channels, self.num_classes, self.dropout_ratio = channels, num_classes, dropout_ratio
            self.conv_cfg, self.norm_cfg, self.act_cfg = conv_cfg, norm_cfg, act_cfg
            self.ignore_index, self.align_corners = ignore_index, align_corners
            self.conv_seg = nn.Conv2d(channels, num_classes, kernel_size=1)
            self.dropout = nn.Dropout2d(dropout_ratio) if dropout_ratio > 0 else None

        def _transform_inputs(self, inputs):
            return [inputs[i] for i in self.in_index]

        def forward_test(self, inputs, img_metas, test_cfg):
            return self.forward(inputs)

        def cls_seg(self, feat):
            if self.dropout is not None:
                feat = self.dropout(feat)
            return self.conv_seg(feat)

    class BaseSegmentor(nn.Module):
        def __init__(self, init_cfg=None):
            super().__init__()

        @property
        def with_neck(self):
            return hasattr(self, "neck") and self.neck is not None

        @property
        def with_auxiliary_head(self):
            return hasattr(self, "auxiliary_head") and self.auxiliary_head is not None

        @property
        def with_decode_head(self):
            return hasattr(self, "decode_head") and self.decode_head is not None

    def resize(input, size=None, scale_factor=None, mode="nearest", align_corners=None, warning=True):
        return F.interpolate(input, size, scale_factor, mode, align_corners)

    _mod("mmseg.models.decode_heads"); _mod("mmseg.models.decode_heads.decode_head", BaseDecodeHead=BaseDecodeHead)
    _mod("mmseg.ops", resize=resize)
    _mod("mmseg.core", add_prefix=lambda d, p: {f"{p}.{k}": v for k, v in d.items()})
    _mod("mmseg.models.segmentors"); _mod("mmseg.models.segmentors.base", BaseSegmentor=BaseSegmentor)
    b = sys.modules["mmseg.models.builder"]
    b.build_backbone = lambda cfg: cfg       # the caller passes constructed modules (see build_segmentor)
    b.build_head = lambda cfg: cfg
    b.build_neck = lambda cfg: cfg
    sys.modules["mmseg.models"].builder = b
    _ns("mmseg_custom.models.decode_heads", os.path.join(SEG, "mmseg_custom/models/decode_heads"))
    _ns("mmseg_custom.models.segmentors", os.path.join(SEG, "mmseg_custom/models/segmentors"))


def build_head(hcfg):
    """The reference's SegformerHead built from the config dict (minus `type`)."""
    install_segmentor()
    from mmseg_custom.models.decode_heads.segformer_head import SegformerHead
    cfg = {k: v for k, v in hcfg.items() if k not in ("type", "loss_decode")}
    with contextlib.redirect_stdout(io.StringIO()):
        return SegformerHead(**cfg).eval()


def build_segmentor(bcfg, hcfg, test_cfg, withcp=False):
    """The reference's EncoderDecoder around the reference backbone + head."""
    install_segmentor()
    from mmseg_custom.models.segmentors.encoder_decoder import EncoderDecoder
    Dict = sys.modules["addict"].Dict
    net = EncoderDecoder(backbone=build_backbone(bcfg, withcp), decode_head=build_head(hcfg), test_cfg=Dict(test_cfg))
    return net.eval()
