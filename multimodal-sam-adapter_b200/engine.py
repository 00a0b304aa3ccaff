"""Inference engine: packs the weights of the parameter containers (nn_modules.py) into kernel-ready
device buffers and runs the MM SAM-Adapter forward as a sequence of this repo's sm_100a kernels.

Data layout in HBM: every activation is channels-last ([B, H, W, C] == token-major [B*H*W, C]), so LayerNorm / LN2d
are row ops, 1x1 convs and Linears are the same GEMM, and the token sequence c = [c2 | c3 | c4] is one
[B, 21*HW/... , C] buffer. Storage is bf16 except the two long RESIDUAL STREAMS, which stay fp32 in HBM: the ViT token
stream x (2 residual adds per block, 24 blocks) and the feature map t of each ConvNeXt tower (36 blocks whose branch is
scaled by a small layer-scale gamma: rounding t to bf16 after every block was measured to be ~80 % of the towers' error
against the fp32 reference, 9.7e-3 -> 2.5e-3 rel-L2 at stage 2 with the stream in fp32). Every tensor-core operand is
bf16 (LayerNorm / depthwise-conv outputs); fp32 is also kept for LN statistics, GEMM / attention accumulation,
MSDeformAttn sampling offsets + attention logits, and the class logits.

Follows the reference forward (segmentation/mmseg_custom/models/backbones/
image_encoder_adapter_bimodal_mix_mod_new_in_twin_convnext_new.py:161-349) step by step; the cited
line ranges are given at each stage below.
"""
import math
import os

import torch

from . import kernels as K
from . import neck as neck_b200


def _on_device(fn):
    """Run a public engine method with the engine's GPU as the current device: kernels launch on the CURRENT device's
    stream and opt in to large shared memory per device, so a model living on cuda:1 must not launch from cuda:0."""
    import functools

    @functools.wraps(fn)
    def wrapped(self, *a, **k):
        with torch.cuda.device(self.dev):
            return fn(self, *a, **k)
    return wrapped


def _bf(t, dev):
    return t.detach().to(device=dev, dtype=torch.bfloat16).contiguous()


def _f32(t, dev):
    return t.detach().to(device=dev, dtype=torch.float32).contiguous()


class U8Input:
    """The network input as the loader delivers it: the two modalities as uint8 HWC frames (K.U8Modality: data + the
    normalisation of the test pipeline) and the padded network input size (H, W). Normalise / pad / HWC -> CHW / patchify
    happen inside the stem and patch-embed operand kernels (mmsam_patchify_u8); the fp32 NCHW tensor never exists."""

    def __init__(self, rgb, aux, hw):
        self.mods = (rgb, aux)
        self.H, self.W = int(hw[0]), int(hw[1])
        self.B = rgb.data.shape[0]
        assert aux.data.shape[0] == self.B and rgb.data.shape[1:3] == aux.data.shape[1:3]

    @property
    def shape_key(self):
        return ("u8", tuple(self.mods[0].data.shape), tuple(self.mods[1].data.shape), self.H, self.W)


class _Lin:
    """weight bf16 [N, K] (K contiguous), bias fp32 [N] or None, optional per-channel scale."""

    def __init__(self, w, b, dev, scale=None):
        self.w = _bf(w.reshape(w.shape[0], -1), dev)
        self.b = None if b is None else _f32(b, dev)
        self.scale = None if scale is None else _f32(scale, dev)
        self.n, self.k = self.w.shape


class _LN:
    def __init__(self, m, dev, eps=None):
        self.w, self.b = _f32(m.weight, dev), _f32(m.bias, dev)
        self.eps = m.eps if eps is None else eps


class _LinLN:
    """LayerNorm -> Linear folded for K.gemm_ln: w = bf16(gamma (.) W), colsum[n] = sum_k w[n,k] (of the ROUNDED weight,
    so that a constant row maps to exactly `b`), b = W beta + bias; the GEMM epilogue applies the row statistics."""

    def __init__(self, lin_w, lin_b, ln, dev):
        W = lin_w.detach().float().reshape(lin_w.shape[0], -1)
        g, beta = ln.weight.detach().float(), ln.bias.detach().float()
        self.w = _bf(W * g[None, :], dev)
        self.colsum = self.w.float().sum(1).contiguous()
        b = W @ beta
        if lin_b is not None:
            b = b + lin_b.detach().float()
        self.b = _f32(b, dev)
        self.eps = ln.eps
        self.n, self.k = self.w.shape


# LayerNorm folded into the consumer GEMM (row statistics + column sums) for the adapter's LN -> Linear pairs; set
# MMSAM_LN_FOLD=0 to run the separate LayerNorm kernel + plain GEMM instead.
LN_FOLD = os.environ.get("MMSAM_LN_FOLD", "1") != "0"
# ConvNeXt block tail (LN + pwconv1 + GELU + pwconv2 + gamma + residual) as one kernel; MMSAM_MLP_FUSED=0: LN + two GEMMs
MLP_FUSED = os.environ.get("MMSAM_MLP_FUSED", "1") != "0"


def _bn_fold(bn, dev):
    s = bn.weight.detach().float() / torch.sqrt(bn.running_var.detach().float() + bn.eps)
    t = bn.bias.detach().float() - bn.running_mean.detach().float() * s
    return _f32(s, dev), _f32(t, dev)


def _pad_rows(w, n):
    if w.shape[0] == n:
        return w
    out = w.new_zeros((n,) + tuple(w.shape[1:]))
    out[: w.shape[0]] = w
    return out


# Shared-memory staged MSDeformAttn kernel (msda_staged2_kernel, profiles/r02_msda_notes.md): 1.27x faster than the
# L1-gather kernel on the extractor (one 64 x 64 value level, 21504 queries), slower on the injector (three levels per
# region leave one CTA per SM): default = extractor only.
MSDA_STAGED = os.environ.get("MMSAM_MSDA_STAGED", "ext")         # "0" | "1" (injector + extractor) | "ext" (extractor only)


class _MSDA:
    """Packed MSDeformAttn weights: one fused query projection (offsets | logits) with fp32 output."""

    def __init__(self, m, dev, query_norm=None, feat_norm=None):
        self.n_heads, self.n_levels, self.n_points = m.n_heads, m.n_levels, m.n_points
        self.value = _Lin(m.value_proj.weight, m.value_proj.bias, dev)
        wq = torch.cat((m.sampling_offsets.weight.detach(), m.attention_weights.weight.detach()), 0)
        bq = torch.cat((m.sampling_offsets.bias.detach(), m.attention_weights.bias.detach()), 0)
        self.qproj = _Lin(wq, bq, dev)
        # the Injector / Extractor feed both projections straight from a LayerNorm: fold it into the GEMMs
        self.value_ln = _LinLN(m.value_proj.weight, m.value_proj.bias, feat_norm, dev) if feat_norm is not None else None
        self.qproj_ln = _LinLN(wq, bq, query_norm, dev) if query_norm is not None else None
        self.out = _Lin(m.output_proj.weight, m.output_proj.bias, dev)
        self.off_bias = m.sampling_offsets.bias.detach().float().cpu()
        self._geoms = {}

    def geom(self, kind, sc):
        """Staged-gather geometry (K.MsdaGeometry): injector = ViT-token queries over the 3 c-levels, extractor =
        the 3 c-grids as queries over the ViT-token map; regions are 8 x 8 (injector: three boxes per CTA) / 8 x 16 (extractor) tiles of the ViT token grid."""
        key = (kind, tuple(sc["s1"]), tuple(sc["s3"]))
        g = self._geoms.get(key)
        if g is None:
            levels, qgrids = (sc["s3"], sc["s1"]) if kind == "inj" else (sc["s1"], sc["s3"])
            g = K.MsdaGeometry(levels, qgrids, sc["s1"][0], (8, 8) if kind == "inj" else (8, 16), self.off_bias, self.n_heads, self.n_levels,
                               self.n_points, margin=int(os.environ.get("MMSAM_MSDA_MARGIN", "2")))
            self._geoms[key] = g
        return g


def pack_block(blk, dev):
    return dict(norm1=_LN(blk.norm1, dev), norm2=_LN(blk.norm2, dev), window=blk.window_size,
                qkv=_Lin(blk.attn.qkv.weight, blk.attn.qkv.bias, dev),
                proj=_Lin(blk.attn.proj.weight, blk.attn.proj.bias, dev),
                lin1=_Lin(blk.mlp.lin1.weight, blk.mlp.lin1.bias, dev),
                lin2=_Lin(blk.mlp.lin2.weight, blk.mlp.lin2.bias, dev),
                rel_h=blk.attn.rel_pos_h.detach().float().to(dev) if blk.attn.use_rel_pos else None,
                rel_w=blk.attn.rel_pos_w.detach().float().to(dev) if blk.attn.use_rel_pos else None)


def pack_interaction(it, dev):
    d = dict(inj=dict(qn=_LN(it.injector.query_norm, dev), fn=_LN(it.injector.feat_norm, dev),
                      attn=_MSDA(it.injector.attn, dev, it.injector.query_norm, it.injector.feat_norm),
                      gamma=_f32(it.injector.gamma, dev)))
    exts = [it.extractor] + (list(it.extra_extractors) if it.extra_extractors is not None else [])
    d["ext"] = []
    for ex in exts:
        e = dict(qn=_LN(ex.query_norm, dev), fn=_LN(ex.feat_norm, dev),
                 attn=_MSDA(ex.attn, dev, ex.query_norm, ex.feat_norm), ffn=None)
        if ex.with_cffn:
            hc = ex.ffn.fc1.weight.shape[0]
            e["ffn"] = dict(norm=_LN(ex.ffn_norm, dev), fc1=_Lin(ex.ffn.fc1.weight, ex.ffn.fc1.bias, dev),
                            fc1_ln=_LinLN(ex.ffn.fc1.weight, ex.ffn.fc1.bias, ex.ffn_norm, dev),
                            dw_w=_f32(ex.ffn.dwconv.dwconv.weight.detach().reshape(hc, 9).t(), dev),
                            dw_b=_f32(ex.ffn.dwconv.dwconv.bias, dev),
                            fc2=_Lin(ex.ffn.fc2.weight, ex.ffn.fc2.bias, dev))
        d["ext"].append(e)
    return d


def pack_head(h, dev):
    """SegformerHead weights (decode_heads/segformer_head.py:24-46): eval BatchNorm folded into the 1x1 convs, the fusion
    conv's weight sliced per input level (see _Ops._head_logits), conv_seg padded to a multiple of 32 classes."""
    hd = {"convs": []}
    for cm in h.convs:
        s, t = _bn_fold(cm.bn, dev)
        w = cm.conv.weight.detach().float().reshape(cm.conv.weight.shape[0], -1).to(dev) * s[:, None]
        hd["convs"].append(_Lin(w, t, dev))
    s, t = _bn_fold(h.fusion_conv.bn, dev)
    w = h.fusion_conv.conv.weight.detach().float().reshape(h.fusion_conv.conv.weight.shape[0], -1).to(dev) * s[:, None]
    chn = w.shape[0]
    hd["fusion_w"] = [w[:, i * chn:(i + 1) * chn].to(torch.bfloat16).contiguous() for i in range(len(h.convs))]
    hd["fusion_shift"] = t.float().contiguous()
    ncls = h.conv_seg.weight.shape[0]
    npad = (ncls + 31) // 32 * 32
    hd["cls"] = _Lin(_pad_rows(h.conv_seg.weight.detach().float().reshape(ncls, -1), npad),
                     _pad_rows(h.conv_seg.bias.detach().float(), npad), dev)
    hd["ncls"], hd["npad"], hd["channels"] = ncls, npad, h.conv_seg.weight.shape[1]
    return hd


def window_maps(B, H, W, ws, C, dev):
    """Row maps of SAM's window_partition / window_unpartition (base/image_encoder.py:504-551)."""
    nwh, nww = (H + ws - 1) // ws, (W + ws - 1) // ws
    b_i = torch.arange(B).view(B, 1, 1)
    y_i = torch.arange(H).view(1, H, 1)
    x_i = torch.arange(W).view(1, 1, W)
    dst = (((b_i * nwh + y_i // ws) * nww + x_i // ws) * (ws * ws) + (y_i % ws) * ws + (x_i % ws)).reshape(-1)
    nrows = B * nwh * nww * ws * ws
    inv = torch.full((nrows,), -1, dtype=torch.int64)
    inv[dst] = torch.arange(B * H * W)
    return dict(win_fwd=dst.to(torch.int32).to(dev), win_inv=inv.to(torch.int32).to(dev), win_rows=nrows,
                win_bp=B * nwh * nww, ws=ws,
                win_buf=torch.zeros((nrows, C), dtype=torch.bfloat16, device=dev))  # pad rows stay 0


def deform_geometry(B, Hi, Wi, dev):
    """deform_inputs (adapter_modules_...new.py:397-431): reference points, level shapes, row maps."""
    s3 = [(Hi // 8, Wi // 8), (Hi // 16, Wi // 16), (Hi // 32, Wi // 32)]
    s1 = [(Hi // 16, Wi // 16)]

    def ref(shapes):
        pts = []
        for (h, w) in shapes:
            ys = torch.linspace(0.5, h - 0.5, h, dtype=torch.float32) / h
            xs = torch.linspace(0.5, w - 0.5, w, dtype=torch.float32) / w
            yy, xx = torch.meshgrid(ys, xs, indexing="ij")
            pts.append(torch.stack((xx.reshape(-1), yy.reshape(-1)), -1))
        return torch.cat(pts, 0).contiguous().to(dev)

    def lv(shapes):
        t = torch.as_tensor(shapes, dtype=torch.long)
        lsi = torch.cat((t.new_zeros((1,)), t.prod(1).cumsum(0)[:-1]))
        return t.to(dev), lsi.to(dev)

    sc = dict(ref1=ref(s1), ref2=ref(s3), lv3=lv(s3), lv1=lv(s1), s1=s1, s3=s3, S3=sum(h * w for h, w in s3))
    offs, o = [], 0
    for (h, w) in s3:
        n = h * w
        rm = (torch.arange(B).view(B, 1) * sc["S3"] + o + torch.arange(n).view(1, n)).reshape(-1)
        offs.append(rm.to(torch.int32).to(dev))
        o += n
    sc["c_rowmaps"] = offs
    return sc


class _Ops:
    """Kernel-level building blocks shared by the full engine and the component runners."""

    def _gemm(self, a, lin, **kw):
        return K.gemm(a, lin.w, bias=lin.b, scale=lin.scale, **kw)

    def _ln(self, x, ln, **kw):
        return K.layernorm(x, ln.w, ln.b, ln.eps, **kw)

    def _gemm_ln(self, x, lnlin, stats_of=None, **kw):
        """lnlin(LayerNorm(x)) with the LayerNorm folded into the GEMM: one statistics pass over x (half the traffic of
        the LayerNorm kernel), no normalised copy of x. stats_of: the shape-cache dict when x is the adapter's token
        buffer `c`: its row statistics are kept there until `c` is written again — the Injector's feat_norm and the
        following Extractor's query_norm normalise the SAME c (only the affine part differs), one pass serves both."""
        st = None
        if stats_of is not None:
            ent = stats_of.get("c_stats")
            if ent is not None and ent[0] == x.data_ptr() and ent[1] == lnlin.eps:
                st = ent[2]
        if st is None:
            st = K.rowstats(x, lnlin.eps)
            if stats_of is not None:
                stats_of["c_stats"] = (x.data_ptr(), lnlin.eps, st)
        return K.gemm_ln(x, lnlin.w, lnlin.b, lnlin.colsum, st, **kw)

    def _proj_ln(self, t, ln, lnlin, lin, stats_of=None, out_dtype=torch.bfloat16):
        """lin(LayerNorm(t)). bf16 rows (the adapter's token buffer c): the LayerNorm is folded into the GEMM (row statistics
        + GEMM on the raw rows). fp32 rows (the ViT token stream): LayerNorm kernel (fp32 -> bf16), then the plain GEMM."""
        if LN_FOLD and lnlin is not None and t.dtype == torch.bfloat16:
            return self._gemm_ln(t, lnlin, stats_of=stats_of, out_dtype=out_dtype)
        return self._gemm(self._ln(t, ln), lin, out_dtype=out_dtype)

    def _msda(self, pk, query, feat, ref, lv, B, geom=None, qn=None, fn=None, c_is=None, sc=None):
        """MSDeformAttn.forward up to (not including) output_proj (ops/modules/ms_deform_attn.py:83-127) on
        query_norm(query) / feat_norm(feat) (adapter_modules_...new.py:494-501, 527-532). c_is: "feat" / "query" tells
        which operand is the adapter's token buffer c (row statistics cached in sc)."""
        value = self._proj_ln(feat, fn, pk.value_ln, pk.value, stats_of=sc if c_is == "feat" else None)           # [B*S, M*D]
        qp = self._proj_ln(query, qn, pk.qproj_ln, pk.qproj, stats_of=sc if c_is == "query" else None,
                           out_dtype=torch.float32)                                                                # [B*Lq, M*L*P*3]
        S = feat.shape[0] // B
        return K.msda_fused(value.view(B, S, -1), lv[0], lv[1], qp, ref, pk.n_heads, pk.n_levels, pk.n_points, geom=geom)

    def _head_logits(self, hd, feats):
        """SegformerHead.forward (decode_heads/segformer_head.py:48-66) on channels-last bf16 maps [B, h, w, C]
        -> fp32 logits [B*h0*w0, npad], (h0, w0)."""
        B, h0, w0, _ = feats[0].shape
        ch = hd["channels"]
        # fusion(concat_i resize(y_i)) = sum_i resize(W_i y_i): every level's slice of the (BN-scaled) fusion weight is
        # applied at the level's own resolution; one kernel adds the up-sampled partial sums, the BN shift and the ReLU
        zs, hws = [], []
        for i, f in enumerate(feats):
            _, h, w, Cf = f.shape
            y = self._gemm(f.reshape(-1, Cf), hd["convs"][i], act="relu")
            zs.append(K.gemm(y, hd["fusion_w"][i]))
            hws.append((h, w))
        if len(zs) > 4:
            raise NotImplementedError("SegformerHead with more than 4 input levels")
        o = K.resize_sum_affine(zs[0], zs[1:], hws[1:], (h0, w0), B, ch, shift=hd["fusion_shift"], relu=True)
        return self._gemm(o, hd["cls"], out_dtype=torch.float32), (h0, w0)

    def _block(self, x, blk, sc, tabs, B, out=None):
        """Block.forward (base/image_encoder.py:382-423). x [B*T, C] (the fp32 token stream) is updated in place unless
        out is given: norm1 / norm2 read it in fp32 and emit the bf16 GEMM operand, the proj / lin2 epilogues add the
        fp32 residual and store fp32."""
        H, W, T = sc["H"], sc["W"], sc["T"]
        if blk["window"] > 0:
            ws = sc["ws"]
            y = self._ln(x, blk["norm1"], out=sc["win_buf"], row_map=sc["win_fwd"])
            qkv = self._gemm(y, blk["qkv"])
            # window_unpartition is fused into the attention store: pad rows are dropped, proj runs on B*T rows
            if ws == 14:       # SAM's window size: un-partition by one TMA store per row tile
                a = K.attention_window(qkv.view(sc["win_bp"], ws * ws, -1), self.nh, B, H, W, tabs[0], tabs[1])
            else:
                a = K.attention(qkv.view(sc["win_bp"], ws * ws, -1), self.nh, (ws, ws), tabs[0], tabs[1],
                                out_map=sc["win_inv"], out_rows=B * T)
            self._gemm(a, blk["proj"], residual=x, out=x)
        else:
            y = self._ln(x, blk["norm1"])
            qkv = self._gemm(y, blk["qkv"])
            a = K.attention(qkv.view(B, T, -1), self.nh, (H, W), tabs[0], tabs[1])
            self._gemm(a.view(-1, self.C), blk["proj"], residual=x, out=x)
        y = self._ln(x, blk["norm2"])
        h = self._gemm(y, blk["lin1"], act="gelu")
        dst = x if out is None else out
        self._gemm(h, blk["lin2"], residual=x, out=dst)
        return dst

    def _injector(self, x, c, inj, sc, B):
        """Injector.forward (adapter_modules_...new.py:525-542); returns a NEW fp32 [B*T, C] buffer (the input
        is one of the saved ViT outputs `outs` and must stay intact)."""
        o = self._msda(inj["attn"], x, c, sc["ref1"], sc["lv3"], B, inj["attn"].geom("inj", sc) if MSDA_STAGED == "1" else None,
                       qn=inj["qn"], fn=inj["fn"], c_is="feat", sc=sc)
        return K.gemm(o.view(-1, o.shape[-1]), inj["attn"].out.w, bias=inj["attn"].out.b, scale=inj["gamma"],
                      residual=x, out=torch.empty_like(x))

    def _extractor(self, c, x, e, sc, B):
        """Extractor.forward (adapter_modules_...new.py:490-511); c [B*S3, C] updated in place."""
        o = self._msda(e["attn"], c, x, sc["ref2"], sc["lv1"], B, e["attn"].geom("ext", sc) if MSDA_STAGED in ("1", "ext") else None,
                       qn=e["qn"], fn=e["fn"], c_is="query", sc=sc)
        sc["c_stats"] = None                                   # c is about to change
        self._gemm(o.view(-1, o.shape[-1]), e["attn"].out, residual=c, out=c)
        f = e["ffn"]
        if f is not None:
            h = self._gemm_ln(c, f["fc1_ln"]) if LN_FOLD else self._gemm(self._ln(c, f["norm"]), f["fc1"])
            hc = h.shape[1]
            h2 = K.dwconv(h, f["dw_w"], f["dw_b"], 3, sc["s3"], B, hc, sc["S3"] * hc, sc["S3"] * hc, act="gelu")
            self._gemm(h2, f["fc2"], residual=c, out=c)


class ComponentRunner(_Ops):
    """Runs a single reference-shaped component (Block / InteractionBlock container) on the kernels;
    used by the component-level parity tests (non-square token grids)."""

    def __init__(self, dim, num_heads, device="cuda"):
        self.C, self.nh, self.dev = dim, num_heads, torch.device(device)

    @_on_device
    @torch.no_grad()
    def block(self, blk_module, x, H, W):
        """x [B, H*W, C] (any float dtype, CUDA) -> same shape, bf16."""
        B = x.shape[0]
        pk = pack_block(blk_module, self.dev)
        sc = dict(H=H, W=W, T=H * W)
        if pk["window"] > 0:
            sc.update(window_maps(B, H, W, pk["window"], self.C, self.dev))
        kh, kw = (pk["window"], pk["window"]) if pk["window"] > 0 else (H, W)
        tabs = (None, None)
        if pk["rel_h"] is not None:
            tabs = (K.relpos_table(pk["rel_h"], kh), K.relpos_table(pk["rel_w"], kw))
        xb = x.reshape(B * H * W, self.C).float().contiguous().clone()
        return self._block(xb, pk, sc, tabs, B).view(B, H * W, self.C)

    @_on_device
    @torch.no_grad()
    def head(self, head_module, feats):
        """SegformerHead.forward on NCHW feature maps (any float dtype, CUDA) -> fp32 logits [B, num_classes, h0, w0]."""
        hd = pack_head(head_module, self.dev)
        nhwc = [f.permute(0, 2, 3, 1).to(torch.bfloat16).contiguous() for f in feats]
        logits, (h0, w0) = self._head_logits(hd, nhwc)
        B = feats[0].shape[0]
        return logits.view(B, h0, w0, -1)[..., : hd["ncls"]].permute(0, 3, 1, 2).contiguous()

    @_on_device
    @torch.no_grad()
    def interaction(self, it_module, x, c, Hi, Wi, blocks=()):
        """InteractionBlock.forward (adapter_modules_...new.py:567-581) for an Hi x Wi image geometry."""
        B = x.shape[0]
        pk = pack_interaction(it_module, self.dev)
        H, W = Hi // 16, Wi // 16
        sc = dict(H=H, W=W, T=H * W)
        sc.update(deform_geometry(B, Hi, Wi, self.dev))
        xb = x.reshape(-1, self.C).float().contiguous().clone()
        cb = c.reshape(-1, self.C).to(torch.bfloat16).contiguous().clone()
        xb = self._injector(xb, cb, pk["inj"], sc, B)
        for blk in blocks:
            xb = self.block(blk, xb.view(B, H * W, self.C), H, W).reshape(-1, self.C)
        for e in pk["ext"]:
            self._extractor(cb, xb, e, sc, B)
        return xb.view(B, H * W, self.C), cb.view(B, -1, self.C)


class EncoderEngine(_Ops):
    def __init__(self, backbone, head=None, device="cuda"):
        self.dev = torch.device(device)
        if self.dev.type != "cuda":
            raise K._lib.MMSamError("EncoderEngine needs a CUDA device: the path has no CPU fallback")
        K._lib.load()
        self.bb = backbone
        self.cfg = backbone.cfg
        self._shape_cache = {}
        self._graphs = {}
        self.graph_launches = 0
        self._pack_backbone(backbone)
        self.head = None
        if head is not None:
            self._pack_head(head)

    # ------------------------------------------------------------------ packing
    def _pack_backbone(self, m):
        dev = self.dev
        self.C = m.embed_dim
        self.nh = m.num_heads
        self.patch = m.patch_size
        self.cin = m.in_ch_im
        pe = m.patch_embed.proj
        self.patch_embed = _Lin(pe.weight, pe.bias, dev)
        self.pos_embed = m.pos_embed.detach().float().to(dev)  # [1, ps, ps, C]
        self.blocks = []
        for blk in m.blocks:
            self.blocks.append(pack_block(blk, dev))
        # --- TwinConvNeXt (base/twin_convnext.py) ---
        tw = m.spm.twin_conv
        self.cnx = {}
        for br in ("x", "y"):
            ds = getattr(tw, f"downsample_layers_{br}")
            st = getattr(tw, f"stages_{br}")
            stages = []
            for i in range(len(tw.depths)):
                e = {}
                if i == 0:
                    e["ds"] = _Lin(ds[0][0].weight, ds[0][0].bias, dev)          # [C0, 3*p*p] (c,ky,kx)
                    e["ds_ln"] = _LN(ds[0][1], dev)
                else:
                    w = ds[i][1].weight.detach()                                    # [Co, Ci, 2, 2]
                    e["ds"] = _Lin(w.permute(0, 2, 3, 1).contiguous(), ds[i][1].bias, dev)  # (ky,kx,ci)
                    e["ds_ln"] = _LN(ds[i][0], dev)
                blocks = []
                for b in st[i]:
                    c = b.depthwise_conv.weight.shape[0]
                    blocks.append(dict(
                        dw_w=_f32(b.depthwise_conv.weight.detach().reshape(c, 49).t(), dev),
                        dw_b=_f32(b.depthwise_conv.bias, dev), ln=_LN(b.norm, dev),
                        pw1=_Lin(b.pointwise_conv1.weight, b.pointwise_conv1.bias, dev),
                        pw2=_Lin(b.pointwise_conv2.weight, b.pointwise_conv2.bias, dev,
                                 scale=b.gamma if b.gamma is not None else None),
                        # LN folded into pwconv1 for the one-launch block tail (mmsam_convnext_mlp_bf16)
                        pw1_ln=_LinLN(b.pointwise_conv1.weight, b.pointwise_conv1.bias, b.norm, dev) if c in (96, 192, 384) else None))
                e["blocks"] = blocks
                e["out_ln"] = _LN(getattr(tw, f"norm_{br}{i}"), dev)
                stages.append(e)
            self.cnx[br] = stages
        self.cnx_channels = list(tw.channels)
        self.neck = neck_b200.NeckB200(m.spm.smart_fusion, dev)
        le = m.level_embed.detach().float()
        self.fc = []
        for i in range(4):
            fc = getattr(m.spm, f"fc{i + 1}")
            b = fc.bias.detach().float()
            if i >= 1:
                b = b + le[i - 1]          # _add_level_embed (..._new.py:149-159) folded into the bias
            self.fc.append(_Lin(fc.weight, b, dev))
        # up: ConvTranspose2d(C, C, 2, 2): weight [Cin, Cout, 2, 2] -> rows (dy, dx, co)
        uw = m.up.weight.detach().permute(2, 3, 1, 0).reshape(4 * self.C, self.C)
        self.up = _Lin(uw, m.up.bias.detach().repeat(4), dev)
        self.inter = []
        for it in m.interactions:
            self.inter.append(pack_interaction(it, dev))
        self.final_bn = [_bn_fold(getattr(m, f"norm{i + 1}"), dev) for i in range(4)]

    def _pack_head(self, h):
        self.head = pack_head(h, self.dev)

    # ------------------------------------------------------------------ shape-dependent constants
    def _shape(self, B, Hi, Wi):
        key = (B, Hi, Wi)
        sc = self._shape_cache.get(key)
        if sc is not None:
            return sc
        dev = self.dev
        p = self.patch
        H, W = Hi // p, Wi // p
        sc = dict(H=H, W=W, T=H * W)
        # pos_embed: always bicubic-resized (..._new.py:136-143)
        pe = torch.nn.functional.interpolate(self.pos_embed.permute(0, 3, 1, 2), size=(H, W), mode="bicubic",
                                             align_corners=False)
        pe = pe.reshape(1, -1, H * W).permute(0, 2, 1)                  # fp32: the residual of the patch-embed GEMM
        sc["pos"] = pe.expand(B, -1, -1).reshape(B * H * W, -1).contiguous()
        ws = max((b["window"] for b in self.blocks), default=0)
        if ws > 0:
            sc.update(window_maps(B, H, W, ws, self.C, dev))
        # relative position tables per block (get_rel_pos, image_encoder.py:554-584)
        tabs = []
        for blk in self.blocks:
            if blk["rel_h"] is None:
                tabs.append((None, None))
            elif blk["window"] > 0:
                tabs.append((K.relpos_table(blk["rel_h"], blk["window"]), K.relpos_table(blk["rel_w"], blk["window"])))
            else:
                tabs.append((K.relpos_table(blk["rel_h"], H), K.relpos_table(blk["rel_w"], W)))
        sc["tabs"] = tabs
        sc.update(deform_geometry(B, Hi, Wi, dev))
        self._shape_cache[key] = sc
        return sc

    # ------------------------------------------------------------------ building blocks
    def _patches(self, img, which, p):
        """bf16 patch rows [(b,py,px), (c,ky,kx)] of modality `which` (0 = rgb, 1 = auxiliary) for a p x p / stride p conv."""
        if isinstance(img, U8Input):
            return K.patchify_u8(img.mods[which], img.H, img.W, p)
        c_off = 0 if which == 0 else self.cin
        return K.patchify(img, c_off, self.cin if which == 0 else img.shape[1] - self.cin, p)

    def _convnext_branch(self, img, c_off, stages, B, Hi, Wi):
        """twin_convnext.py:445-476 for one modality; returns the 4 normalised stage outputs bf16 [B*h*w, C]. The tower's
        feature map t is an fp32 residual stream (see the module docstring); dwconv / LN read it in fp32 and emit bf16."""
        feats = []
        t = None
        h, w = Hi, Wi
        f32 = torch.float32
        for i, st in enumerate(stages):
            if i == 0:
                pch = self._patches(img, 0 if c_off == 0 else 1, 4)
                h, w = Hi // 4, Wi // 4
                t = self._gemm(pch, st["ds"], out_dtype=f32)
                t = self._ln(t, st["ds_ln"], out=t, out_dtype=f32)
            else:
                pt = self._ln(t, st["ds_ln"], patchify_hw=(h, w))
                h, w = h // 2, w // 2
                t = self._gemm(pt, st["ds"], out_dtype=f32)
            C = t.shape[1]
            for b in st["blocks"]:      # ConvNeXtBlock (twin_convnext.py:98-132)
                y = K.dwconv(t, b["dw_w"], b["dw_b"], 7, [(h, w)], B, C, h * w * C, h * w * C)
                f = b["pw1_ln"]
                if f is not None and MLP_FUSED:
                    # norm -> pwconv1 -> GELU -> pwconv2 -> gamma -> += t in one launch (the 4C intermediate stays on chip)
                    K.convnext_mlp(y, f.w, f.colsum, f.b, b["pw2"].w, b["pw2"].b, b["pw2"].scale, t, f.eps)
                    continue
                y = self._ln(y, b["ln"])
                y = self._gemm(y, b["pw1"], act="gelu")
                self._gemm(y, b["pw2"], residual=t, out=t)
            feats.append((self._ln(t, st["out_ln"]), h, w))
        return feats

    # ------------------------------------------------------------------ forward
    @_on_device
    @torch.no_grad()
    def backbone_nhwc(self, img, debug=None):
        """img fp32 [B, 3+3, Hi, Wi] on the device, or a U8Input (uint8 HWC frames + normalisation)
        -> [f1, f2, f3, f4] channels-last bf16 [B, h, w, C]."""
        if isinstance(img, U8Input):
            B, Hi, Wi = img.B, img.H, img.W
            if any(m.data.device != self.dev for m in img.mods):
                raise K._lib.MMSamError(f"input frames not on {self.dev}")
        else:
            if not img.is_cuda:
                raise K._lib.MMSamError("input must be a CUDA tensor")
            if img.device != self.dev:
                raise K._lib.MMSamError(f"input on {img.device}, engine packed for {self.dev}")
            img = img.contiguous().float()
            B, _, Hi, Wi = img.shape
        if Hi % 32 or Wi % 32:
            raise ValueError("input height/width must be multiples of 32")
        isz = self.cfg.get("img_size")
        if isz is not None and (Hi, Wi) != (isz, isz):
            # the reference's GFFM.norm = nn.LayerNorm(H/4 * W/4) ... is sized from img_size at construction
            # (adapter_modules_...new.py:236-241, 347-354) and raises on any other input size; so do we
            raise K._lib.MMSamError(f"the model is resolution-locked to img_size={isz} (fusion-neck LayerNorms over H*W); "
                                    f"got a {Hi} x {Wi} input")
        sc = self._shape(B, Hi, Wi)
        H, W, T, C, S3 = sc["H"], sc["W"], sc["T"], self.C, sc["S3"]
        # --- SPM (adapter_modules_...new.py:929-964): twin ConvNeXt -> fusion neck -> fc1..4 ---
        # The two modality towers are independent until the fusion neck: the second one runs on a side stream (a parallel
        # branch of the CUDA graph), so each kernel's partially filled last wave is covered by the other tower's work.
        if os.environ.get("MMSAM_TOWER_STREAMS", "1") != "0":
            main = torch.cuda.current_stream()
            if getattr(self, "_tower_stream", None) is None:
                self._tower_stream = torch.cuda.Stream(device=self.dev)
            side = self._tower_stream
            side.wait_stream(main)
            with torch.cuda.stream(side):
                fy = self._convnext_branch(img, self.cin, self.cnx["y"], B, Hi, Wi)
            fx = self._convnext_branch(img, 0, self.cnx["x"], B, Hi, Wi)
            main.wait_stream(side)
        else:
            fx = self._convnext_branch(img, 0, self.cnx["x"], B, Hi, Wi)
            fy = self._convnext_branch(img, self.cin, self.cnx["y"], B, Hi, Wi)
        nd = [] if debug is not None else None
        fused = self.neck(fx, fy, B, debug=nd)                          # 4 x [B*h*w, 2*Ci] bf16
        if debug is not None:
            debug["neck"] = nd
        if debug is not None:
            debug.update(fx=[(t.clone(), h, w) for t, h, w in fx], fy=[(t.clone(), h, w) for t, h, w in fy], fused=[t.clone() for t in fused])
        c1 = self._gemm(fused[0], self.fc[0])                            # [B*16T, C]
        c = torch.empty((B * S3, C), dtype=torch.bfloat16, device=self.dev)
        sc["c_stats"] = None                                    # row statistics of c cached by _gemm_ln: none yet
        for i in range(3):
            self._gemm(fused[i + 1], self.fc[i + 1], out=c, row_map=sc["c_rowmaps"][i], out_rows=B * S3)
        # --- patch embed + pos embed (image_encoder.py:662-671, ..._new.py:268-278) ---
        pch = self._patches(img, 0, self.patch)
        x = self._gemm(pch, self.patch_embed, residual=sc["pos"], out_dtype=torch.float32)     # fp32 token stream
        if debug is not None:
            debug.update(c1=c1.clone(), c_0=c.clone(), x_0=x.clone())
        # --- interactions (..._new.py:283-292; adapter_modules_...new.py:567-581) ---
        outs = []
        idxs = self.cfg["interaction_indexes"]
        for i, (lo, hi) in enumerate(idxs):
            it = self.inter[i]
            x = self._injector(x, c, it["inj"], sc, B)
            for bi in range(lo, hi + 1):
                self._block(x, self.blocks[bi], sc, sc["tabs"][bi], B)
            for e in it["ext"]:
                self._extractor(c, x, e, sc, B)
            outs.append(x)
            if debug is not None:
                debug[f"x{i}"] = x.clone()
                debug[f"c_{i}"] = c.clone()
        # --- tail (..._new.py:316-337): up(c2) + c1, add resized ViT features, eval BatchNorm ---
        c3d = c.view(B, S3, C)
        n2, n3 = 4 * T, T
        g1 = torch.empty((B, 16 * T, C), dtype=torch.bfloat16, device=self.dev)
        c1v = c1.view(B, 16 * T, C)
        for b in range(B):
            K.gemm(c3d[b, :n2], self.up.w, bias=self.up.b, residual=c1v[b], out=g1[b], pixel_shuffle=(2 * H, 2 * W))
        fs = []
        bases = [(g1, 16 * T * C, 0), (c3d, S3 * C, 0), (c3d, S3 * C, n2), (c3d, S3 * C, n2 + n3)]
        sizes = [(4 * H, 4 * W), (2 * H, 2 * W), (H, W), (H // 2, W // 2)]
        add_vit = self.cfg.get("add_vit_feature", True)
        for i in range(4):
            base, bstride, roff = bases[i]
            basev = base[:, roff:] if roff else base
            s, t = self.final_bn[i]
            if add_vit:
                f = K.resize_add_affine(outs[i], (H, W), sizes[i], B, C, base=basev, scale=s, shift=t,
                                        base_bstride=bstride)
            else:
                f = K.resize_add_affine(basev, sizes[i], sizes[i], B, C, scale=s, shift=t, src_bstride=bstride)
            fs.append(f)
        return fs

    @_on_device
    @torch.no_grad()
    def head_logits(self, feats):
        """SegformerHead.forward (decode_heads/segformer_head.py:48-66) -> fp32 logits [B*h*w, npad], (h, w)."""
        return self._head_logits(self.head, feats)

    @staticmethod
    def _in_hw(img):
        return (img.H, img.W) if isinstance(img, U8Input) else tuple(img.shape[2:])

    @_on_device
    @torch.no_grad()
    def segment(self, img, out_hw=None, crop_hw=None):
        """encode_decode_test + whole_inference_dim(_cut) + softmax/argmax -> uint8 labels [B, H, W]
        (segmentors/encoder_decoder.py:96-117, 329-414, 417-508). The head logits go to the label map in one kernel when
        out_hw is the input size (the second resize of whole_inference_dim is then the identity); otherwise the logits are
        first resized to the input size (encode_decode), then to out_hw, as the reference does."""
        Hi, Wi = self._in_hw(img)
        feats = self.backbone_nhwc(img)
        B = feats[0].shape[0]
        logits, (h0, w0) = self.head_logits(feats)
        out_hw = (Hi, Wi) if out_hw is None else tuple(out_hw)
        if out_hw != (Hi, Wi):
            logits, (h0, w0) = K.resize_logits(logits, B, (h0, w0), (Hi, Wi)), (Hi, Wi)
        return K.upsample_argmax(logits, B, (h0, w0), self.head["ncls"], out_hw, crop_hw)

    @_on_device
    @torch.no_grad()
    def segment_graphed(self, img, out_hw=None, crop_hw=None):
        """segment() replayed from a CUDA graph (one graph per input shape, captured on first use after an eager
        warm-up): the ~4000 kernel launches of a ViT-L forward are submitted with one cudaGraphLaunch, so the
        host is off the critical path. The returned label tensor is the graph's static output buffer (valid
        until the next call with the same shape)."""
        u8 = isinstance(img, U8Input)
        key = (img.shape_key if u8 else tuple(img.shape), None if out_hw is None else tuple(out_hw),
               None if crop_hw is None else tuple(crop_hw))
        ent = self._graphs.get(key)
        if ent is None:
            if len(self._graphs) >= 4:
                self._graphs.pop(next(iter(self._graphs)))
            if u8:
                mods = [K.U8Modality(m.data.detach().clone().contiguous(), m.mean, m.std, m.prescale, m.to_rgb, m.pad_val)
                        for m in img.mods]
                static_in = U8Input(mods[0], mods[1], (img.H, img.W))
            else:
                static_in = img.detach().clone().float().contiguous()
            self.segment(static_in, out_hw, crop_hw)          # warm-up: shape caches, kernel attributes
            torch.cuda.synchronize()
            n0 = K.LAUNCHES
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                out = self.segment(static_in, out_hw, crop_hw)
            ent = (graph, static_in, out, K.LAUNCHES - n0)
            self._graphs[key] = ent
        graph, static_in, out, nl = ent
        if u8:
            for dst, src in zip(static_in.mods, img.mods):
                if (dst.mean, dst.std, dst.prescale, dst.to_rgb, dst.pad_val) != (src.mean, src.std, src.prescale, src.to_rgb, src.pad_val):
                    raise K._lib.MMSamError("normalisation changed since this input shape was captured: call invalidate()")
                dst.data.copy_(src.data, non_blocking=True)
        else:
            static_in.copy_(img, non_blocking=True)
        graph.replay()
        self.graph_launches += nl
        return out
