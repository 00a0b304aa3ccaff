"""Oracle: MM SAM-Adapter encoder + Segformer head forward on CPU (fp32/fp64). TEST INFRASTRUCTURE.

A functional restatement (state-dict in, tensors out) of the reference forward pass. Each function
cites the reference lines it follows (paths relative to /root/reference/segmentation/):

  B   = mmseg_custom/models/backbones/image_encoder_adapter_bimodal_mix_mod_new_in_twin_convnext_new.py
  V   = mmseg_custom/models/backbones/base/image_encoder.py
  A   = mmseg_custom/models/backbones/adapter_modules_multimodal_mix_mod_new_in_twin_convnext_new.py
  T   = mmseg_custom/models/backbones/base/twin_convnext.py
  D   = ops/modules/ms_deform_attn.py
  H   = mmseg_custom/models/decode_heads/segformer_head.py
  E   = mmseg_custom/models/segmentors/encoder_decoder.py

Pinned against the reference's own modules (imported in the build container through
tools/ref_shim.py) by tools/make_golden.py; the committed fixtures are replayed by
tests/test_oracle_golden.py. The head follows mmseg 0.20.2 BaseDecodeHead / mmcv ConvModule
semantics as restated in SURVEY.md §8(c) (those packages are not vendored in the reference:
parity unpinned at that boundary).
"""
import math

import torch
import torch.nn.functional as F

from .msda import ms_deform_attn_core

ARCH = {  # T:187-228
    "atto": ([2, 2, 6, 2], [40, 80, 160, 320]), "femto": ([2, 2, 6, 2], [48, 96, 192, 384]),
    "pico": ([2, 2, 6, 2], [64, 128, 256, 512]), "nano": ([2, 2, 8, 2], [80, 160, 320, 640]),
    "tiny": ([3, 3, 9, 3], [96, 192, 384, 768]), "small": ([3, 3, 27, 3], [96, 192, 384, 768]),
    "base": ([3, 3, 27, 3], [128, 256, 512, 1024]), "large": ([3, 3, 27, 3], [192, 384, 768, 1536]),
    "xlarge": ([3, 3, 27, 3], [256, 512, 1024, 2048]), "huge": ([3, 3, 27, 3], [352, 704, 1408, 2816]),
}


def arch_of(cfg):
    a = cfg.get("arch", "base")
    if isinstance(a, str):
        return ARCH[a]
    return list(a["depths"]), list(a["channels"])


class SD:
    """state-dict view with a key prefix."""

    def __init__(self, sd, prefix=""):
        self.sd, self.p = sd, prefix

    def __call__(self, k):
        return self.sd[self.p + k]

    def sub(self, k):
        return SD(self.sd, self.p + k + ".")

    def has(self, k):
        return (self.p + k) in self.sd


def ln(x, s, eps):
    return F.layer_norm(x, (x.shape[-1],), s("weight"), s("bias"), eps)


def linear(x, s):
    return F.linear(x, s("weight"), s("bias") if s.has("bias") else None)


# ------------------------------------------------------------------------------------------------
# SAM ViT block  (V:331-623)
# ------------------------------------------------------------------------------------------------
def get_rel_pos(q_size, k_size, rel_pos):
    """V:554-584 (linear interpolation of the table, float coords truncated by .long())."""
    max_rel_dist = int(2 * max(q_size, k_size) - 1)
    if rel_pos.shape[0] != max_rel_dist:
        r = F.interpolate(rel_pos.reshape(1, rel_pos.shape[0], -1).permute(0, 2, 1), size=max_rel_dist, mode="linear")
        r = r.reshape(-1, max_rel_dist).permute(1, 0)
    else:
        r = rel_pos
    qc = torch.arange(q_size)[:, None] * max(k_size / q_size, 1.0)
    kc = torch.arange(k_size)[None, :] * max(q_size / k_size, 1.0)
    rc = (qc - kc) + (k_size - 1) * max(q_size / k_size, 1.0)
    return r[rc.long()]


def attention_core(q, k, v, h, w, rel_pos_h=None, rel_pos_w=None):
    """V:492-498, 587-623. q,k,v [Bn, h*w, hd] -> [Bn, h*w, hd]; rel-pos bias uses the UNSCALED q."""
    Bn, T, hd = q.shape
    attn = (q * hd ** -0.5) @ k.transpose(-2, -1)
    if rel_pos_h is not None:
        Rh = get_rel_pos(h, h, rel_pos_h)
        Rw = get_rel_pos(w, w, rel_pos_w)
        rq = q.reshape(Bn, h, w, hd)
        rel_h = torch.einsum("bhwc,hkc->bhwk", rq, Rh)
        rel_w = torch.einsum("bhwc,wkc->bhwk", rq, Rw)
        attn = (attn.view(-1, h, w, h, w) + rel_h[:, :, :, :, None] + rel_w[:, :, :, None, :]).view(-1, h * w, h * w)
    return attn.softmax(dim=-1) @ v


def attention(x, s, num_heads):
    """V:483-501. x [B',h,w,C]."""
    Bp, h, w, C = x.shape
    hd = C // num_heads
    qkv = linear(x, s.sub("qkv")).reshape(Bp, h * w, 3, num_heads, hd).permute(2, 0, 3, 1, 4)
    q, k, v = qkv.reshape(3, Bp * num_heads, h * w, hd).unbind(0)
    o = attention_core(q, k, v, h, w, s("rel_pos_h") if s.has("rel_pos_h") else None,
                       s("rel_pos_w") if s.has("rel_pos_w") else None)
    o = o.view(Bp, num_heads, h, w, hd).permute(0, 2, 3, 1, 4).reshape(Bp, h, w, C)
    return linear(o, s.sub("proj"))


def vit_block(x, s, H, W, window, num_heads):
    """V:382-423 with window_partition / window_unpartition V:504-551 (pad AFTER norm1, zeros)."""
    B, N, C = x.shape
    x = x.view(B, H, W, C)
    shortcut = x
    y = ln(x, s.sub("norm1"), 1e-6)
    if window > 0:
        ph, pw = (window - H % window) % window, (window - W % window) % window
        y = F.pad(y, (0, 0, 0, pw, 0, ph))
        Hp, Wp = H + ph, W + pw
        y = y.view(B, Hp // window, window, Wp // window, window, C).permute(0, 1, 3, 2, 4, 5).reshape(-1, window, window, C)
    y = attention(y, s.sub("attn"), num_heads)
    if window > 0:
        y = y.view(B, Hp // window, Wp // window, window, window, C).permute(0, 1, 3, 2, 4, 5).reshape(B, Hp, Wp, C)
        y = y[:, :H, :W]
    x = shortcut + y
    m = s.sub("mlp")
    x = x + linear(F.gelu(linear(ln(x, s.sub("norm2"), 1e-6), m.sub("lin1"))), m.sub("lin2"))
    return x.reshape(B, N, C)


# ------------------------------------------------------------------------------------------------
# MSDeformAttn, Injector, Extractor  (D:83-130, A:412-431, 474-581)
# ------------------------------------------------------------------------------------------------
def reference_points(shapes, dtype):
    """A:397-409."""
    pts = []
    for (h, w) in shapes:
        ys = torch.linspace(0.5, h - 0.5, h, dtype=dtype) / h
        xs = torch.linspace(0.5, w - 0.5, w, dtype=dtype) / w
        yy, xx = torch.meshgrid(ys, xs, indexing="ij")
        pts.append(torch.stack((xx.reshape(-1), yy.reshape(-1)), -1))
    return torch.cat(pts, 0)[None, :, None]  # [1, sum, 1, 2]


def deform_inputs(h, w, dtype):
    """A:412-431. Returns (ref1, shapes1), (ref2, shapes2) for an input image of h x w."""
    s3 = [(h // 8, w // 8), (h // 16, w // 16), (h // 32, w // 32)]
    s1 = [(h // 16, w // 16)]
    return (reference_points(s1, dtype), s3), (reference_points(s3, dtype), s1)


def msdeform_attn(query, ref, feat, shapes, s, n_heads, n_points, core=ms_deform_attn_core):
    """D:83-130 (2-d reference points branch)."""
    N, Lq, C = query.shape
    L = len(shapes)
    value = linear(feat, s.sub("value_proj"))
    value = value.view(N, feat.shape[1], n_heads, value.shape[-1] // n_heads)
    off = linear(query, s.sub("sampling_offsets")).view(N, Lq, n_heads, L, n_points, 2)
    aw = linear(query, s.sub("attention_weights")).view(N, Lq, n_heads, L * n_points)
    aw = F.softmax(aw, -1).view(N, Lq, n_heads, L, n_points)
    norm = torch.tensor([[w, h] for (h, w) in shapes], dtype=query.dtype)
    loc = ref[:, :, None, :, None, :] + off / norm[None, None, None, :, None, :]
    out = core(value, shapes, loc, aw)
    return linear(out, s.sub("output_proj"))


def injector(x, c, ref, shapes, s, n_heads, n_points):
    """A:525-542."""
    a = msdeform_attn(ln(x, s.sub("query_norm"), 1e-6), ref, ln(c, s.sub("feat_norm"), 1e-6), shapes, s.sub("attn"), n_heads, n_points)
    return x + s("gamma") * a


def conv_ffn(x, s, H, W):
    """A:446-471: fc1 -> shared 3x3 depthwise conv over the 3 token grids -> GELU -> fc2."""
    B, N, _ = x.shape
    x = linear(x, s.sub("fc1"))
    C = x.shape[-1]
    n = N // 21
    dw, db = s("dwconv.dwconv.weight"), s("dwconv.dwconv.bias")
    outs = []
    for lo, hi, (h, w) in ((0, 16 * n, (2 * H, 2 * W)), (16 * n, 20 * n, (H, W)), (20 * n, N, (H // 2, W // 2))):
        t = x[:, lo:hi].transpose(1, 2).reshape(B, C, h, w)
        outs.append(F.conv2d(t, dw, db, padding=1, groups=C).flatten(2).transpose(1, 2))
    return linear(F.gelu(torch.cat(outs, 1)), s.sub("fc2"))


def extractor(c, x, ref, shapes, s, n_heads, n_points, H, W):
    """A:490-511."""
    a = msdeform_attn(ln(c, s.sub("query_norm"), 1e-6), ref, ln(x, s.sub("feat_norm"), 1e-6), shapes, s.sub("attn"), n_heads, n_points)
    c = c + a
    return c + conv_ffn(ln(c, s.sub("ffn_norm"), 1e-6), s.sub("ffn"), H, W)


# ------------------------------------------------------------------------------------------------
# TwinConvNeXt  (T:98-132, 445-476)
# ------------------------------------------------------------------------------------------------
def ln2d(x, s, eps=1e-6):
    return F.layer_norm(x.permute(0, 2, 3, 1), (x.shape[1],), s("weight"), s("bias"), eps).permute(0, 3, 1, 2)


def convnext_block(x, s):
    C = x.shape[1]
    y = F.conv2d(x, s("depthwise_conv.weight"), s("depthwise_conv.bias"), padding=3, groups=C)
    y = y.permute(0, 2, 3, 1)
    y = F.layer_norm(y, (C,), s("norm.weight"), s("norm.bias"), 1e-6)
    y = linear(F.gelu(linear(y, s.sub("pointwise_conv1"))), s.sub("pointwise_conv2"))
    if s.has("gamma"):
        y = y * s("gamma")
    return x + y.permute(0, 3, 1, 2)


def twin_convnext(x, y, s, depths):
    outs = []
    for br, t in (("x", x), ("y", y)):
        feats = []
        for i, d in enumerate(depths):
            ds = s.sub(f"downsample_layers_{br}.{i}")
            if i == 0:
                t = F.conv2d(t, ds("0.weight"), ds("0.bias"), stride=ds("0.weight").shape[-1])
                t = ln2d(t, ds.sub("1"))
            else:
                t = ln2d(t, ds.sub("0"))
                t = F.conv2d(t, ds("1.weight"), ds("1.bias"), stride=2)
            for j in range(d):
                t = convnext_block(t, s.sub(f"stages_{br}.{i}.{j}"))
            feats.append(ln2d(t, s.sub(f"norm_{br}{i}")))
        outs.append(feats)
    return [torch.cat((a, b), 1) for a, b in zip(*outs)]


# ------------------------------------------------------------------------------------------------
# RoadFormer2Neck fusion  (A:39-394)
# ------------------------------------------------------------------------------------------------
def gfe(x, s, heads=8, groups=32):
    """A:75-145: WithBias LN over channels (eps 1e-5) -> channel attention -> x + attn(...)."""
    b, c, h, w = x.shape
    t = x.flatten(2).transpose(1, 2)
    mu = t.mean(-1, keepdim=True)
    var = t.var(-1, keepdim=True, unbiased=False)
    t = (t - mu) / torch.sqrt(var + 1e-5) * s("norm1.body.weight") + s("norm1.body.bias")
    n = t.transpose(1, 2).reshape(b, c, h, w)
    a = s.sub("attn")
    qkv = F.conv2d(F.conv2d(n, a("qkv1.weight"), None, groups=groups), a("qkv2.weight"), None, padding=1, groups=groups)
    q, k, v = qkv.chunk(3, dim=1)
    q, k, v = (z.reshape(b, heads, c // heads, h * w) for z in (q, k, v))
    q, k = F.normalize(q, dim=-1), F.normalize(k, dim=-1)
    att = ((q @ k.transpose(-2, -1)) * a("scale")).softmax(-1)
    o = (att @ v).reshape(b, c, h, w)
    o = F.conv2d(o, a("proj.weight"), None)
    att_out = n + o * a("scale2")          # AttentionBase returns x_in + out*scale2 where x_in = norm1(x)
    return x + att_out


def mobilenet_v2(x, s):
    """A:281-295."""
    y = F.relu6(F.conv2d(x, s("bottleneckBlock.0.weight")))
    y = F.relu6(F.conv2d(y, s("bottleneckBlock.2.weight"), None, padding=1, groups=y.shape[1]))
    y = F.conv2d(y, s("bottleneckBlock.4.weight"))
    return y * s("scale") + x


def gffm(x, s):
    """A:234-267: cross-modal C x C attention, LayerNorm over the flattened spatial axis."""
    b, c2, h, w = x.shape
    c = c2 // 2
    fx, fy = x[:, :c].reshape(b, c, -1), x[:, c:].reshape(b, c, -1)
    ax = F.softmax(torch.bmm(fx, fy.transpose(1, 2)), -1)
    ay = F.softmax(torch.bmm(fy, fx.transpose(1, 2)), -1)
    ox = torch.bmm(ax, fy) * s("gammax.scale") + fx
    oy = torch.bmm(ay, fx) * s("gammay.scale") + fy
    o = torch.cat((ox, oy), 1)
    o = F.layer_norm(o, (h * w,), s("norm.weight"), s("norm.bias"), 1e-5)
    return o.view(b, c2, h, w)


def gated_mlp(x, s):
    """A:110-132 (ffn_expansion_factor = 1)."""
    y = F.conv2d(x, s("project_in.weight"))
    y = F.conv2d(y, s("dwconv.weight"), None, padding=1, groups=y.shape[1] // 2)
    a, g = y.chunk(2, 1)
    return F.conv2d(F.gelu(a) * g, s("project_out.weight"))


def ffrm(x, s):
    """A:148-162: GAP -> 1x1 conv (no bias) -> GN(32) -> ReLU (ConvModule default act) -> sigmoid gate."""
    a = F.adaptive_avg_pool2d(x, 1)
    a = F.conv2d(a, s("conv_atten.conv.weight"))
    a = F.relu(F.group_norm(a, 32, s("conv_atten.gn.weight"), s("conv_atten.gn.bias"), 1e-5))
    return x + x * torch.sigmoid(a)


def coord_att(x, s):
    """A:176-221 (CA = x + CoordinateAttention(x)); bn1 in eval mode."""
    n, c, h, w = x.shape
    xh = x.mean(3, keepdim=True)                      # [n,c,h,1]
    xw = x.mean(2, keepdim=True).permute(0, 1, 3, 2)  # [n,c,w,1]
    y = torch.cat((xh, xw), 2)
    y = F.conv2d(y, s("conv1.weight"), s("conv1.bias"))
    y = F.batch_norm(y, s("bn1.running_mean"), s("bn1.running_var"), s("bn1.weight"), s("bn1.bias"), False, 0.0, 1e-5)
    y = y * F.relu6(y + 3) / 6
    yh, yw = torch.split(y, [h, w], 2)
    ah = torch.sigmoid(F.conv2d(yh, s("conv_h.weight"), s("conv_h.bias")))
    aw = torch.sigmoid(F.conv2d(yw.permute(0, 1, 3, 2), s("conv_w.weight"), s("conv_w.bias")))
    return x + x * aw * ah


def roadformer_neck(feats, s, stages=None):
    """A:364-394."""
    out = []
    for i, f in enumerate(feats):
        c = f.shape[1] // 2
        rgb, aux = f[:, :c], f[:, c:]
        g = torch.cat((gfe(rgb, s.sub(f"global_feature_encoder_rgb.{i}")), gfe(aux, s.sub(f"global_feature_encoder_sne.{i}"))), 1)
        l = torch.cat((mobilenet_v2(rgb, s.sub(f"local_feature_encoder_rgb.{i}")), mobilenet_v2(aux, s.sub(f"local_feature_encoder_sne.{i}"))), 1)
        if stages is not None:
            stages.append(dict(g=g.clone(), l=l.clone()))
        g = gffm(g, s.sub(f"fuse_blocks.{i}"))
        l = gated_mlp(l, s.sub(f"detail_feature_extractions.{i}"))
        if stages is not None:
            stages[-1].update(o_ln=g.clone(), lo=l.clone())
        g = ffrm(g, s.sub(f"enhance_blocks.{i}"))
        sc = s.sub(f"scale_layers.{i}")
        f2 = g * sc("scale1") + l * sc("scale2")
        if stages is not None:
            stages[-1].update(f=f2.clone())
        out.append(coord_att(f2, s.sub(f"ca_blocks.{i}.coord_atten")))
    return out


def spm_bimodal(x, y, s, depths, stages=None):
    """A:929-964."""
    tw = twin_convnext(x, y, s.sub("twin_conv"), depths)
    if stages is not None:
        stages["twin"] = [t.clone() for t in tw]
    nst = [] if stages is not None else None
    feats = roadformer_neck(tw, s.sub("smart_fusion"), nst)
    if stages is not None:
        stages["fused"] = [t.clone() for t in feats]
        stages["neck"] = nst
    cs = []
    for i, f in enumerate(feats):
        t = F.conv2d(f, s(f"fc{i + 1}.weight"), s(f"fc{i + 1}.bias"))
        cs.append(t.flatten(2).transpose(1, 2))
    return cs


# ------------------------------------------------------------------------------------------------
# Backbone  (B:161-349)
# ------------------------------------------------------------------------------------------------
def backbone_forward(sd, cfg, img, prefix="", stages=None):
    """img [B, 3+3, H, W] -> [f1, f2, f3, f4] (NCHW). `stages` (dict) collects intermediates."""
    s = SD(sd, prefix)
    depths, _ = arch_of(cfg)
    nh, dh, npts = cfg["num_heads"], cfg["deform_num_heads"], cfg.get("n_points", 4)
    win, glob = cfg.get("window_size", 14), list(cfg.get("global_attn_indexes", [5, 11, 17, 23]))
    cin = cfg["modalities_ch"][cfg["modalities_name"].index("rgb")]
    x, xo = img[:, :cin], img[:, cin:]
    B, _, Hi, Wi = x.shape
    c1, c2, c3, c4 = spm_bimodal(x, xo, s.sub("spm"), depths, stages)
    if stages is not None:
        stages.update(c1=c1, c2=c2, c3=c3, c4=c4)
    le = s("level_embed")
    c = torch.cat((c2 + le[0], c3 + le[1], c4 + le[2]), 1)
    (ref1, shapes1), (ref2, shapes2) = deform_inputs(Hi, Wi, x.dtype)
    pw = s("patch_embed.proj.weight")
    t = F.conv2d(x, pw, s("patch_embed.proj.bias"), stride=pw.shape[-1])
    H, W = t.shape[2], t.shape[3]
    t = t.permute(0, 2, 3, 1).flatten(1, 2)
    pe = F.interpolate(s("pos_embed").permute(0, 3, 1, 2), size=(H, W), mode="bicubic", align_corners=False)
    t = t + pe.reshape(1, -1, H * W).permute(0, 2, 1)
    if stages is not None:
        stages["x_0"] = t.clone()
        stages["c_0"] = c.clone()
    C = t.shape[-1]
    outs = []
    idxs = cfg["interaction_indexes"]
    for i, (lo, hi) in enumerate(idxs):
        it = s.sub(f"interactions.{i}")
        t = injector(t, c, ref1, shapes1, it.sub("injector"), dh, npts)
        for b in range(lo, hi + 1):
            t = vit_block(t, s.sub(f"blocks.{b}"), H, W, 0 if b in glob else win, nh)
        c = extractor(c, t, ref2, shapes2, it.sub("extractor"), dh, npts, H, W)
        if i == len(idxs) - 1 and cfg.get("use_extra_extractor", True):
            for j in range(2):
                c = extractor(c, t, ref2, shapes2, it.sub(f"extra_extractors.{j}"), dh, npts, H, W)
        outs.append(t.transpose(1, 2).reshape(B, C, H, W))
        if stages is not None:
            stages[f"x{i}"] = t
            stages[f"c_{i}"] = c
    n2, n3 = c2.shape[1], c3.shape[1]
    c1 = c1.transpose(1, 2).reshape(B, C, 4 * H, 4 * W)
    c2 = c[:, :n2].transpose(1, 2).reshape(B, C, 2 * H, 2 * W)
    c3 = c[:, n2:n2 + n3].transpose(1, 2).reshape(B, C, H, W)
    c4 = c[:, n2 + n3:].transpose(1, 2).reshape(B, C, H // 2, W // 2)
    c1 = F.conv_transpose2d(c2, s("up.weight"), s("up.bias"), stride=2) + c1
    if cfg.get("add_vit_feature", True):
        x1, x2, x3, x4 = outs
        c1 = c1 + F.interpolate(x1, scale_factor=4, mode="bilinear", align_corners=False)
        c2 = c2 + F.interpolate(x2, scale_factor=2, mode="bilinear", align_corners=False)
        c3 = c3 + x3
        c4 = c4 + F.interpolate(x4, scale_factor=0.5, mode="bilinear", align_corners=False)
    fs = []
    for i, f in enumerate((c1, c2, c3, c4)):
        n = s.sub(f"norm{i + 1}")
        fs.append(F.batch_norm(f, n("running_mean"), n("running_var"), n("weight"), n("bias"), False, 0.0, 1e-5))
    return fs


# ------------------------------------------------------------------------------------------------
# Segformer head + inference post-processing  (H:48-66, E:96-117, 329-414, 417-508)
# ------------------------------------------------------------------------------------------------
def _conv_bn_relu(x, s):
    y = F.conv2d(x, s("conv.weight"))
    y = F.batch_norm(y, s("bn.running_mean"), s("bn.running_var"), s("bn.weight"), s("bn.bias"), False, 0.0, 1e-5)
    return F.relu(y)


def segformer_head(sd, feats, prefix="decode_head."):
    s = SD(sd, prefix)
    size = feats[0].shape[2:]
    outs = [F.interpolate(_conv_bn_relu(f, s.sub(f"convs.{i}")), size=size, mode="bilinear", align_corners=False)
            for i, f in enumerate(feats)]
    o = _conv_bn_relu(torch.cat(outs, 1), s.sub("fusion_conv"))
    return F.conv2d(o, s("conv_seg.weight"), s("conv_seg.bias"))  # Dropout2d is identity in eval


def segmentor_logits(sd, cfg, img, test_cfg=None):
    """encode_decode_test + whole_inference_dim / _dim_cut (E:96-117, 329-414): logits at image size."""
    feats = backbone_forward(sd, cfg, img, prefix="backbone.")
    logits = segformer_head(sd, feats)
    logits = F.interpolate(logits, size=img.shape[2:], mode="bilinear", align_corners=False)
    if test_cfg:
        if test_cfg.get("dim") is not None:
            logits = F.interpolate(logits, size=tuple(test_cfg["dim"]), mode="bilinear", align_corners=False)
        if test_cfg.get("cut_dim") is not None:
            cw, ch = test_cfg["cut_dim"]
            logits = logits[:, :, :ch, :cw]
    return logits


def encode_decode(sd, cfg, img):
    """E:85-95 / 96-117: head logits resized (bilinear, align_corners=False) to the input size."""
    feats = backbone_forward(sd, cfg, img, prefix="backbone.")
    return F.interpolate(segformer_head(sd, feats), size=img.shape[2:], mode="bilinear", align_corners=False)


def slide_inference(sd, cfg, img, test_cfg, rescale=False, ori_shape=None):
    """E:191-234: overlapping crops of the network input size, zero-padded logits summed and divided by the count."""
    h_stride, w_stride = test_cfg["stride"]
    h_crop, w_crop = test_cfg["crop_size"]
    B, _, h_img, w_img = img.shape
    h_grids = max(h_img - h_crop + h_stride - 1, 0) // h_stride + 1
    w_grids = max(w_img - w_crop + w_stride - 1, 0) // w_stride + 1
    preds = count = None
    for hi in range(h_grids):
        for wi in range(w_grids):
            y1, x1 = hi * h_stride, wi * w_stride
            y2, x2 = min(y1 + h_crop, h_img), min(x1 + w_crop, w_img)
            y1, x1 = max(y2 - h_crop, 0), max(x2 - w_crop, 0)
            lg = encode_decode(sd, cfg, img[:, :, y1:y2, x1:x2])
            if preds is None:
                preds = img.new_zeros((B, lg.shape[1], h_img, w_img))
                count = img.new_zeros((B, 1, h_img, w_img))
            preds = preds + F.pad(lg, (x1, w_img - x2, y1, h_img - y2))
            count[:, :, y1:y2, x1:x2] += 1
    preds = preds / count
    if rescale:
        preds = F.interpolate(preds, size=tuple(ori_shape[:2]), mode="bilinear", align_corners=False)
    return preds


def inference(sd, cfg, img, test_cfg, rescale=True, ori_shape=None, flip=False, flip_direction="horizontal"):
    """E:417-469: mode dispatch (slide / whole / whole_dim / whole_dim_cut), softmax, flip. Returns probabilities."""
    mode = test_cfg.get("mode", "whole")
    ori_shape = tuple(img.shape[2:]) if ori_shape is None else tuple(ori_shape)
    if mode == "slide":
        lg = slide_inference(sd, cfg, img, test_cfg, rescale, ori_shape)
    else:
        lg = encode_decode(sd, cfg, img)
        if mode == "whole":                      # E:310-327
            if rescale:
                lg = F.interpolate(lg, size=ori_shape[:2], mode="bilinear", align_corners=False)
        elif mode == "whole_dim":                # E:329-362 (returns nothing when rescale is False)
            if not rescale:
                raise ValueError("whole_inference_dim returns None with rescale=False (encoder_decoder.py:333-362)")
            lg = F.interpolate(lg, size=tuple(test_cfg["dim"]), mode="bilinear", align_corners=False)
        elif mode == "whole_dim_cut":            # E:364-391: the backbone returns (feats, None) -> the tuple branch
            if rescale:
                lg = F.interpolate(lg, size=tuple(test_cfg["dim"]), mode="bilinear", align_corners=False)
            cw, ch = test_cfg["cut_dim"]
            lg = lg[:, :, :ch, :cw]
        else:
            raise ValueError(mode)
    out = F.softmax(lg, dim=1)
    if flip:
        out = out.flip(dims=(3,)) if flip_direction == "horizontal" else out.flip(dims=(2,))
    return out


def simple_test(sd, cfg, img, test_cfg=None, rescale=True, ori_shape=None, flip=False, flip_direction="horizontal"):
    """E:471-508: softmax then argmax -> int64 labels [B,H,W]. With test_cfg=None / no "mode": the whole_dim(_cut) logits
    of segmentor_logits (kept for the earlier tests)."""
    if test_cfg is not None and "mode" in test_cfg:
        return inference(sd, cfg, img, test_cfg, rescale, ori_shape, flip, flip_direction).argmax(dim=1)
    return F.softmax(segmentor_logits(sd, cfg, img, test_cfg), dim=1).argmax(dim=1)
