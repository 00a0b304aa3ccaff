"""Parameter containers with the reference's state-dict schema.

These nn.Modules hold the weights under exactly the names/shapes of the reference modules, so the
released `backbone.*` / `decode_head.*` checkpoints load with load_state_dict, and reproduce the
reference's random init. They contain NO forward arithmetic: all compute is done by
engine.EncoderEngine through the sm_100a kernels.

Reference structure followed (paths under segmentation/mmseg_custom/models/):
  backbones/base/image_encoder.py:187-328 (ImageEncoderViT), :331-423 (Block), :426-501 (Attention)
  backbones/adapter_modules_multimodal_mix_mod_new_in_twin_convnext_new.py:75-394, 434-581, 861-964
  backbones/base/twin_convnext.py:23-132, 134-380
  backbones/image_encoder_adapter_bimodal_mix_mod_new_in_twin_convnext_new.py:29-147
  decode_heads/segformer_head.py:11-46
"""
import math

import torch
import torch.nn as nn

ARCH_SETTINGS = {  # base/twin_convnext.py:187-228
    "atto": ([2, 2, 6, 2], [40, 80, 160, 320]), "femto": ([2, 2, 6, 2], [48, 96, 192, 384]),
    "pico": ([2, 2, 6, 2], [64, 128, 256, 512]), "nano": ([2, 2, 8, 2], [80, 160, 320, 640]),
    "tiny": ([3, 3, 9, 3], [96, 192, 384, 768]), "small": ([3, 3, 27, 3], [96, 192, 384, 768]),
    "base": ([3, 3, 27, 3], [128, 256, 512, 1024]), "large": ([3, 3, 27, 3], [192, 384, 768, 1536]),
    "xlarge": ([3, 3, 27, 3], [256, 512, 1024, 2048]), "huge": ([3, 3, 27, 3], [352, 704, 1408, 2816]),
}


def _ln(dim, eps=1e-6):
    return nn.LayerNorm(dim, eps=eps)


class _Holder(nn.Module):
    def forward(self, *a, **k):  # pragma: no cover
        raise RuntimeError("parameter container: compute goes through mmsam_b200.engine")


# ---------------------------------------------------------------- SAM ViT -------------------------
class MLPBlock(_Holder):
    def __init__(self, dim, mlp_dim):
        super().__init__()
        self.lin1 = nn.Linear(dim, mlp_dim)
        self.lin2 = nn.Linear(mlp_dim, dim)


class Attention(_Holder):
    def __init__(self, dim, num_heads, qkv_bias, use_rel_pos, input_size):
        super().__init__()
        self.num_heads = num_heads
        hd = dim // num_heads
        self.qkv = nn.Linear(dim, dim * 3, bias=qkv_bias)
        self.proj = nn.Linear(dim, dim)
        self.use_rel_pos = use_rel_pos
        if use_rel_pos:
            self.rel_pos_h = nn.Parameter(torch.zeros(2 * input_size[0] - 1, hd))
            self.rel_pos_w = nn.Parameter(torch.zeros(2 * input_size[1] - 1, hd))


class Block(_Holder):
    def __init__(self, dim, num_heads, mlp_ratio, qkv_bias, use_rel_pos, window_size, input_size):
        super().__init__()
        self.norm1 = _ln(dim)
        self.attn = Attention(dim, num_heads, qkv_bias, use_rel_pos,
                              input_size if window_size == 0 else (window_size, window_size))
        self.norm2 = _ln(dim)
        self.mlp = MLPBlock(dim, int(dim * mlp_ratio))
        self.window_size = window_size


class PatchEmbed(_Holder):
    def __init__(self, kernel_size, stride, in_chans, embed_dim):
        super().__init__()
        self.proj = nn.Conv2d(in_chans, embed_dim, kernel_size=kernel_size, stride=stride)


# ---------------------------------------------------------------- MSDeformAttn params ------------
class MSDeformAttnParams(_Holder):
    """Weights of ops.modules.MSDeformAttn (ops/modules/ms_deform_attn.py:28-81); the callable public
    module lives in mmsam_b200.ops.modules and subclasses this."""

    def __init__(self, d_model=256, n_levels=4, n_heads=8, n_points=4, ratio=1.0):
        super().__init__()
        if d_model % n_heads != 0:
            raise ValueError("d_model must be divisible by n_heads, but got {} and {}".format(d_model, n_heads))
        self.im2col_step = 64
        self.d_model, self.n_levels, self.n_heads, self.n_points, self.ratio = d_model, n_levels, n_heads, n_points, ratio
        self.sampling_offsets = nn.Linear(d_model, n_heads * n_levels * n_points * 2)
        self.attention_weights = nn.Linear(d_model, n_heads * n_levels * n_points)
        self.value_proj = nn.Linear(d_model, int(d_model * ratio))
        self.output_proj = nn.Linear(int(d_model * ratio), d_model)
        self._reset_parameters()

    def _reset_parameters(self):
        nn.init.constant_(self.sampling_offsets.weight.data, 0.)
        thetas = torch.arange(self.n_heads, dtype=torch.float32) * (2.0 * math.pi / self.n_heads)
        grid = torch.stack([thetas.cos(), thetas.sin()], -1)
        grid = (grid / grid.abs().max(-1, keepdim=True)[0]).view(self.n_heads, 1, 1, 2).repeat(1, self.n_levels, self.n_points, 1)
        for i in range(self.n_points):
            grid[:, :, i, :] *= i + 1
        with torch.no_grad():
            self.sampling_offsets.bias = nn.Parameter(grid.view(-1))
        nn.init.constant_(self.attention_weights.weight.data, 0.)
        nn.init.constant_(self.attention_weights.bias.data, 0.)
        nn.init.xavier_uniform_(self.value_proj.weight.data)
        nn.init.constant_(self.value_proj.bias.data, 0.)
        nn.init.xavier_uniform_(self.output_proj.weight.data)
        nn.init.constant_(self.output_proj.bias.data, 0.)


# ---------------------------------------------------------------- adapter -------------------------
class DWConv(_Holder):
    def __init__(self, dim):
        super().__init__()
        self.dwconv = nn.Conv2d(dim, dim, 3, 1, 1, bias=True, groups=dim)


class ConvFFN(_Holder):
    def __init__(self, in_features, hidden_features):
        super().__init__()
        self.fc1 = nn.Linear(in_features, hidden_features)
        self.dwconv = DWConv(hidden_features)
        self.fc2 = nn.Linear(hidden_features, in_features)


class Extractor(_Holder):
    def __init__(self, dim, num_heads, n_points, n_levels, deform_ratio, with_cffn, cffn_ratio, attn_cls):
        super().__init__()
        self.query_norm = _ln(dim)
        self.feat_norm = _ln(dim)
        self.attn = attn_cls(d_model=dim, n_levels=n_levels, n_heads=num_heads, n_points=n_points, ratio=deform_ratio)
        self.with_cffn = with_cffn
        if with_cffn:
            self.ffn = ConvFFN(dim, int(dim * cffn_ratio))
            self.ffn_norm = _ln(dim)


class Injector(_Holder):
    def __init__(self, dim, num_heads, n_points, n_levels, deform_ratio, init_values, attn_cls):
        super().__init__()
        self.query_norm = _ln(dim)
        self.feat_norm = _ln(dim)
        self.attn = attn_cls(d_model=dim, n_levels=n_levels, n_heads=num_heads, n_points=n_points, ratio=deform_ratio)
        self.gamma = nn.Parameter(init_values * torch.ones(dim))


class InteractionBlock(_Holder):
    def __init__(self, dim, num_heads, n_points, with_cffn, cffn_ratio, init_values, deform_ratio, extra_extractor, attn_cls):
        super().__init__()
        self.injector = Injector(dim, num_heads, n_points, 3, deform_ratio, init_values, attn_cls)
        self.extractor = Extractor(dim, num_heads, n_points, 1, deform_ratio, with_cffn, cffn_ratio, attn_cls)
        if extra_extractor:
            self.extra_extractors = nn.Sequential(*[
                Extractor(dim, num_heads, n_points, 1, deform_ratio, with_cffn, cffn_ratio, attn_cls) for _ in range(2)])
        else:
            self.extra_extractors = None


# ---------------------------------------------------------------- TwinConvNeXt --------------------
class ConvNeXtBlock(_Holder):
    def __init__(self, c, layer_scale_init_value):
        super().__init__()
        self.depthwise_conv = nn.Conv2d(c, c, kernel_size=7, padding=3, groups=c)
        self.norm = _ln(c)
        self.pointwise_conv1 = nn.Linear(c, 4 * c)
        self.pointwise_conv2 = nn.Linear(4 * c, c)
        self.gamma = nn.Parameter(layer_scale_init_value * torch.ones(c)) if layer_scale_init_value > 0 else None


class TwinConvNeXt(_Holder):
    def __init__(self, arch="tiny", in_channels=3, stem_patch_size=4, layer_scale_init_value=1e-6, **_unused):
        super().__init__()
        if isinstance(arch, str):
            depths, channels = ARCH_SETTINGS[arch]
        else:
            depths, channels = list(arch["depths"]), list(arch["channels"])
        self.depths, self.channels, self.stem_patch_size = depths, channels, stem_patch_size
        for br in ("x", "y"):
            ds = nn.ModuleList()
            ds.append(nn.Sequential(nn.Conv2d(in_channels, channels[0], kernel_size=stem_patch_size, stride=stem_patch_size),
                                    _ln(channels[0])))
            stages = nn.ModuleList()
            for i, (d, c) in enumerate(zip(depths, channels)):
                if i >= 1:
                    ds.append(nn.Sequential(_ln(channels[i - 1]), nn.Conv2d(channels[i - 1], c, kernel_size=2, stride=2)))
                stages.append(nn.Sequential(*[ConvNeXtBlock(c, layer_scale_init_value) for _ in range(d)]))
            setattr(self, f"downsample_layers_{br}", ds)
            setattr(self, f"stages_{br}", stages)
        for i, c in enumerate(channels):  # registration order follows the reference (x_i, y_i interleaved)
            self.add_module(f"norm_x{i}", _ln(c))
            self.add_module(f"norm_y{i}", _ln(c))


# ---------------------------------------------------------------- fusion neck ---------------------
class _Body(_Holder):
    def __init__(self, dim):
        super().__init__()
        self.weight = nn.Parameter(torch.ones(dim))
        self.bias = nn.Parameter(torch.zeros(dim))


class _NeckLN(_Holder):
    def __init__(self, dim):
        super().__init__()
        self.body = _Body(dim)


class AttentionBase(_Holder):
    def __init__(self, dim, num_heads=8, groups=32):
        super().__init__()
        self.num_heads = num_heads
        self.scale = nn.Parameter(torch.ones(num_heads, 1, 1))
        self.scale2 = nn.Parameter(torch.tensor(1.0))
        self.qkv1 = nn.Conv2d(dim, dim * 3, kernel_size=1, groups=groups, bias=False)
        self.qkv2 = nn.Conv2d(dim * 3, dim * 3, kernel_size=3, padding=1, groups=groups, bias=False)
        self.proj = nn.Conv2d(dim, dim, kernel_size=1, bias=False)


class GFE(_Holder):
    def __init__(self, dim):
        super().__init__()
        self.norm1 = _NeckLN(dim)
        self.attn = AttentionBase(dim)


class MobileNetV2(_Holder):
    def __init__(self, c):
        super().__init__()
        h = c * 2
        self.bottleneckBlock = nn.Sequential(
            nn.Conv2d(c, h, 1, bias=False), nn.ReLU6(inplace=True),
            nn.Conv2d(h, h, 3, stride=1, padding=1, groups=h, bias=False), nn.ReLU6(inplace=True),
            nn.Conv2d(h, c, 1, bias=False))
        self.scale = nn.Parameter(torch.tensor(0.0))


class _Scale(_Holder):
    def __init__(self, v):
        super().__init__()
        self.scale = nn.Parameter(torch.tensor(float(v)))


class GFFM(_Holder):
    def __init__(self, feat_scale):
        super().__init__()
        self.gammax = _Scale(0)
        self.gammay = _Scale(0)
        self.norm = nn.LayerNorm(feat_scale[0] * feat_scale[1])


class Scale2(_Holder):
    def __init__(self):
        super().__init__()
        self.scale1 = nn.Parameter(torch.tensor(1.0))
        self.scale2 = nn.Parameter(torch.tensor(1.0))


class GatedMlp(_Holder):
    def __init__(self, c):
        super().__init__()
        self.project_in = nn.Conv2d(c, 2 * c, kernel_size=1, bias=False)
        self.dwconv = nn.Conv2d(2 * c, 2 * c, kernel_size=3, stride=1, padding=1, groups=c, bias=False)
        self.project_out = nn.Conv2d(c, c, kernel_size=1, bias=False)


class _ConvGN(_Holder):
    def __init__(self, c):
        super().__init__()
        self.conv = nn.Conv2d(c, c, 1, bias=False)
        self.gn = nn.GroupNorm(32, c)


class FFRM(_Holder):
    def __init__(self, c):
        super().__init__()
        self.conv_atten = _ConvGN(c)
        nn.init.kaiming_uniform_(self.conv_atten.conv.weight, a=1, mode="fan_in", nonlinearity="leaky_relu")


class CoordinateAttention(_Holder):
    def __init__(self, c, reduction=32):
        super().__init__()
        mip = max(8, c // reduction)
        self.conv1 = nn.Conv2d(c, mip, 1)
        self.bn1 = nn.BatchNorm2d(mip)
        self.conv_h = nn.Conv2d(mip, c, 1)
        self.conv_w = nn.Conv2d(mip, c, 1)
        for m in (self.conv1, self.conv_h, self.conv_w):
            nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu")
            nn.init.constant_(m.bias, 0)


class CA(_Holder):
    def __init__(self, c):
        super().__init__()
        self.coord_atten = CoordinateAttention(c)


class RoadFormer2Neck(_Holder):
    def __init__(self, in_channels, img_scale):
        super().__init__()
        self.in_channels = in_channels
        self.enhance_blocks = nn.ModuleList([FFRM(c) for c in in_channels])
        self.global_feature_encoder_rgb = nn.ModuleList([GFE(c // 2) for c in in_channels])
        self.global_feature_encoder_sne = nn.ModuleList([GFE(c // 2) for c in in_channels])
        self.local_feature_encoder_rgb = nn.ModuleList([MobileNetV2(c // 2) for c in in_channels])
        self.local_feature_encoder_sne = nn.ModuleList([MobileNetV2(c // 2) for c in in_channels])
        self.ca_blocks = nn.ModuleList([CA(c) for c in in_channels])
        scales = [(img_scale[0] // 2 ** (i + 2), img_scale[1] // 2 ** (i + 2)) for i in range(len(in_channels))]
        self.fuse_blocks = nn.ModuleList([GFFM(s) for s in scales])
        self.scale_layers = nn.ModuleList([Scale2() for _ in in_channels])
        self.detail_feature_extractions = nn.ModuleList([GatedMlp(c) for c in in_channels])


def _conv_init(m):
    fan_out = m.kernel_size[0] * m.kernel_size[1] * m.out_channels // m.groups
    m.weight.data.normal_(0, math.sqrt(2.0 / fan_out))
    if m.bias is not None:
        m.bias.data.zero_()


class SpatialPriorModuleBimodal(_Holder):
    def __init__(self, inplanes, embed_dim, img_size, arch):
        super().__init__()
        self.twin_conv = TwinConvNeXt(arch=arch, in_channels=3, stem_patch_size=4, layer_scale_init_value=1.0)
        chans = [4 * inplanes, 8 * inplanes, 16 * inplanes, 32 * inplanes]
        for i, c in enumerate(chans):
            fc = nn.Conv2d(c, embed_dim, kernel_size=1, bias=True)
            _conv_init(fc)
            setattr(self, f"fc{i + 1}", fc)
        if isinstance(img_size, int):
            img_size = (img_size, img_size)
        self.smart_fusion = RoadFormer2Neck(chans, img_size)
        tw = self.twin_conv.channels
        if [2 * c for c in tw] != chans:
            raise ValueError(f"conv_inplane={inplanes} is inconsistent with ConvNeXt channels {tw} "
                             "(need 2*channels[i] == 4*conv_inplane*2**i)")


class SegformerHeadParams(_Holder):
    """decode_heads/segformer_head.py:11-46 + mmseg BaseDecodeHead.conv_seg."""

    def __init__(self, in_channels, channels, num_classes):
        super().__init__()
        self.convs = nn.ModuleList()
        for c in in_channels:
            m = _Holder()
            m.conv = nn.Conv2d(c, channels, 1, bias=False)
            m.bn = nn.BatchNorm2d(channels)
            self.convs.append(m)
        self.fusion_conv = _Holder()
        self.fusion_conv.conv = nn.Conv2d(channels * len(in_channels), channels, 1, bias=False)
        self.fusion_conv.bn = nn.BatchNorm2d(channels)
        self.conv_seg = nn.Conv2d(channels, num_classes, kernel_size=1)
        nn.init.normal_(self.conv_seg.weight, mean=0, std=0.01)
        nn.init.constant_(self.conv_seg.bias, 0)
