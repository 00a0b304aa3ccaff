"""RoadFormer2Neck fusion (adapter_modules_multimodal_mix_mod_new_in_twin_convnext_new.py:75-394) on the
sm_100a kernels. Per pyramid level (C = both modalities, ci = C/2, HW pixels, channels-last bf16 tokens):

  GFE (per modality)   LN(eps 1e-5) [layernorm, dual output n and x+n] -> qkv1 grouped 1x1 as a block-diagonal
                       GEMM -> qkv2 grouped 3x3 [conv3x3, tcgen05 implicit GEMM] -> q k^T over HW + row norms
                       [gram] -> softmax(cos * temperature) and proj folded into one [ci, ci] matrix per image
                       -> v . Weff^T + (x + n) [gemm, per image]
  MobileNetV2          1x1 + ReLU6 [gemm] -> depthwise 3x3 + ReLU6 [dwconv] -> 1x1 * scale + x [gemm]
  GFFM                 cross-modal energy fx^T fy over HW [gram] -> row softmaxes -> attn . f + f [gemm, per
                       image] -> LayerNorm over the SPATIAL axis: column statistics [colstats]
  Mlp                  1x1 [gemm] -> grouped 3x3 (2 ch / group) [conv3x3] -> gelu(a) * b [gate] -> 1x1 [gemm]
  FFRM / Scale2 / CA   GAP from the statistics, C x C + GroupNorm gate, weighted sum, coordinate-attention
                       pools [combine_pool] -> two tiny 1x1s -> out = f * (1 + a_w a_h) [ca_apply]

Everything is one of this repo's kernels, including the O(B*C^2)-sized steps between the pixel passes (softmax of the
[ch, ch] / [ci, ci] matrices with proj folded into Weff [gfe_weff], the GFFM softmaxes [gffm_softmax], LayerNorm-over-HW
statistics + FFRM gate [ffrm_gate], coordinate-attention vectors [ca_vectors]). Every reduction adds its partial sums in
a fixed order: no floating-point atomics, bit-identical results run to run. torch is used for buffers only.
"""
import os

import torch

from . import kernels as K


def _f32(t, dev):
    return t.detach().to(device=dev, dtype=torch.float32).contiguous()


def _bf(t, dev):
    return t.detach().to(device=dev, dtype=torch.bfloat16).contiguous()


def _blockdiag_1x1(w, groups):
    """Conv2d(k=1, groups) weight [Cout, Cin/groups, 1, 1] -> dense [Cout, Cin] with zeros off the groups."""
    Cout, cgi = w.shape[0], w.shape[1]
    Cin, cgo = cgi * groups, Cout // groups
    out = torch.zeros((Cout, Cin), dtype=torch.float32)
    wf = w.detach().float().cpu().reshape(Cout, cgi)
    for g in range(groups):
        out[g * cgo:(g + 1) * cgo, g * cgi:(g + 1) * cgi] = wf[g * cgo:(g + 1) * cgo]
    return out


class NeckB200:
    def __init__(self, m, dev):
        self.dev = dev
        self.levels = []
        self._side = []
        for i, C in enumerate(m.in_channels):
            ci = C // 2
            lv = dict(C=C, ci=ci)
            for name, key in (("rgb", "global_feature_encoder_rgb"), ("sne", "global_feature_encoder_sne")):
                g = getattr(m, key)[i]
                a = g.attn
                heads, groups = a.num_heads, a.qkv1.groups
                lv["gfe_" + name] = dict(
                    lnw=_f32(g.norm1.body.weight, dev), lnb=_f32(g.norm1.body.bias, dev),
                    w1=_bf(_blockdiag_1x1(a.qkv1.weight, groups), dev),
                    w2=K.pack_conv3x3_weight(a.qkv2.weight.to(dev), groups), groups=groups, heads=heads,
                    temp=_f32(a.scale.reshape(-1), dev), scale2=_f32(a.scale2.reshape(1), dev),
                    wp=_f32(a.proj.weight.reshape(ci, ci), dev))
            for name, key in (("rgb", "local_feature_encoder_rgb"), ("sne", "local_feature_encoder_sne")):
                l = getattr(m, key)[i]
                bb = l.bottleneckBlock
                lv["mb_" + name] = dict(
                    w0=_bf(bb[0].weight.reshape(2 * ci, ci), dev),
                    dw=_f32(bb[2].weight.detach().reshape(2 * ci, 9).t(), dev),
                    w4=_bf(bb[4].weight.reshape(ci, 2 * ci), dev),
                    scale=_f32(l.scale.detach().reshape(1).expand(ci), dev))
            f = m.fuse_blocks[i]
            wpix, bpix = _f32(f.norm.weight, dev), _f32(f.norm.bias, dev)
            lv["gffm"] = dict(gx=_f32(f.gammax.scale.detach().reshape(1).expand(ci), dev),
                              gy=_f32(f.gammay.scale.detach().reshape(1).expand(ci), dev), wpix=wpix, bpix=bpix,
                              sum_w=float(wpix.double().sum()), mean_b=float(bpix.double().mean()), eps=f.norm.eps)
            d = m.detail_feature_extractions[i]
            lv["mlp"] = dict(win=_bf(d.project_in.weight.reshape(2 * C, C), dev),
                             wdw=K.pack_conv3x3_weight(d.dwconv.weight.to(dev), d.dwconv.groups), groups=d.dwconv.groups,
                             wout=_bf(d.project_out.weight.reshape(C, C), dev))
            e = m.enhance_blocks[i].conv_atten
            lv["ffrm"] = dict(w=_f32(e.conv.weight.reshape(C, C), dev), gw=_f32(e.gn.weight, dev), gb=_f32(e.gn.bias, dev),
                              groups=e.gn.num_groups, eps=e.gn.eps)
            s = m.scale_layers[i]
            lv["s1"], lv["s2"] = float(s.scale1.detach()), float(s.scale2.detach())
            ca = m.ca_blocks[i].coord_atten
            bn = ca.bn1
            bs = bn.weight.detach().float() / torch.sqrt(bn.running_var.detach().float() + bn.eps)
            bt = bn.bias.detach().float() - bn.running_mean.detach().float() * bs
            mip = ca.conv1.weight.shape[0]
            lv["ca"] = dict(w1=_f32(ca.conv1.weight.reshape(mip, C), dev), b1=_f32(ca.conv1.bias, dev), bs=_f32(bs, dev),
                            bt=_f32(bt, dev), wh=_f32(ca.conv_h.weight.reshape(C, mip), dev), bh=_f32(ca.conv_h.bias, dev),
                            ww=_f32(ca.conv_w.weight.reshape(C, mip), dev), bw=_f32(ca.conv_w.bias, dev))
            self.levels.append(lv)

    # ------------------------------------------------------------------ pieces
    @staticmethod
    def _per_image_gemm(a, w, rows, **kw):
        """a [B*rows, K] x w[b] (w bf16 [B, N, K]) per image: one grouped launch when a 256-row tile cannot straddle two
        images, else one launch per image (maps whose pixel count is not a multiple of 256: FMB's 200^2 ... 25^2)."""
        B = w.shape[0]
        out, res = kw.pop("out"), kw.pop("residual")
        if rows % 256 == 0:
            return K.gemm_grouped(a, w, rows, residual=res, out=out, **kw)
        for b in range(B):
            sl = slice(b * rows, (b + 1) * rows)
            K.gemm(a[sl], w[b], residual=res[sl], out=out[sl], **kw)
        return out

    def _gfe(self, x, p, g_out, col0, B, h, w, ci):
        """GFE.forward (:133-145) + AttentionBase.forward (:90-109); writes g_out[:, col0:col0+ci]."""
        HW = h * w
        r = torch.empty_like(x)
        n = K.layernorm(x, p["lnw"], p["lnb"], 1e-5, out2=r)
        qa = K.gemm(n, p["w1"])
        qkv = K.conv3x3(qa, p["w2"], B, h, w, 3 * ci, 3 * ci, p["groups"])
        heads = p["heads"]
        S, nq, nk = K.gram(qkv, 3 * ci, 0, ci, ci, B, HW, blk=ci // heads, norms=True)
        # per-head cosine similarity * temperature -> softmax, with proj (and scale2) folded into one [ci, ci] matrix / image
        weff = K.gfe_weff(S, nq, nk, p["temp"], p["wp"], p["scale2"], heads)
        self._per_image_gemm(qkv[:, 2 * ci:], weff, HW, residual=r, out=g_out[:, col0:col0 + ci])

    def _mobilenet(self, x, p, l_out, col0, B, h, w, ci):
        """MobileNetV2.forward (:293-295); writes l_out[:, col0:col0+ci]."""
        h1 = K.gemm(x, p["w0"], act="relu6")
        h2 = K.dwconv(h1, p["dw"], None, 3, [(h, w)], B, 2 * ci, h * w * 2 * ci, h * w * 2 * ci, act="relu6")
        K.gemm(h2, p["w4"], scale=p["scale"], residual=x, out=l_out[:, col0:col0 + ci])

    def _level(self, lv, tx, ty, B, h, w, debug=None):
        dev = self.dev
        C, ci, HW = lv["C"], lv["ci"], h * w
        gf = lv["gffm"]
        if gf["wpix"].numel() != HW:
            raise K._lib.MMSamError(f"fusion neck built for {gf['wpix'].numel()} pixels at this level, got {h} x {w}: the "
                                    "model is resolution-locked to img_size (GFFM.norm = LayerNorm(H*W))")
        g = torch.empty((B * HW, C), dtype=torch.bfloat16, device=dev)
        l = torch.empty((B * HW, C), dtype=torch.bfloat16, device=dev)
        self._gfe(tx, lv["gfe_rgb"], g, 0, B, h, w, ci)
        self._gfe(ty, lv["gfe_sne"], g, ci, B, h, w, ci)
        self._mobilenet(tx, lv["mb_rgb"], l, 0, B, h, w, ci)
        self._mobilenet(ty, lv["mb_sne"], l, ci, B, h, w, ci)
        # ---- GFFM (:242-267) ----
        ax, ay = K.gffm_softmax(K.gram(g, C, 0, ci, ci, B, HW, blk=0))
        o = torch.empty_like(g)
        self._per_image_gemm(g[:, ci:], ax, HW, scale=gf["gx"], residual=g[:, :ci], out=o[:, :ci])
        self._per_image_gemm(g[:, :ci], ay, HW, scale=gf["gy"], residual=g[:, ci:], out=o[:, ci:])
        # ---- LayerNorm over HW (statistics only) + FFRM gate (:148-162): 1x1 conv -> GN(32) -> ReLU -> sigmoid ----
        ff = lv["ffrm"]
        mu, rstd, gate_v = K.ffrm_gate(K.colstats_part(o, gf["wpix"], B, HW, C), HW, gf["sum_w"], gf["mean_b"], gf["eps"],
                                       ff["w"], ff["gw"], ff["gb"], ff["groups"], ff["eps"])
        # ---- gated Mlp on the local branch (:110-132) ----
        mp = lv["mlp"]
        a1 = K.gemm(l, mp["win"])
        a2 = K.conv3x3(a1, mp["wdw"], B, h, w, 2 * C, 2 * C, mp["groups"])
        u = K.gate(a2, C)
        lo = K.gemm(u, mp["wout"])
        # ---- LN_HW * gate * s1 + local * s2, coordinate-attention pools and vectors ----
        f, ph, pwp = K.combine_pool(o, lo, mu, rstd, gate_v, gf["wpix"], gf["bpix"], lv["s1"], lv["s2"], B, h, w, C)
        if debug is not None:
            debug.append(dict(g=g.clone(), l=l.clone(), o=o.clone(), lo=lo.clone(), f=f.clone(), mu=mu.clone(), rstd=rstd.clone(),
                              gate=gate_v.clone()))
        ca = lv["ca"]
        ah, aw = K.ca_vectors(ph, pwp, B, h, w, C, ca["w1"], ca["b1"], ca["bs"], ca["bt"], ca["wh"], ca["bh"], ca["ww"], ca["bw"])
        return K.ca_apply(f, ah, aw, B, h, w, C)

    @torch.no_grad()
    def __call__(self, fx, fy, B, debug=None):
        """fx / fy: per level (tokens bf16 [B*h*w, ci], h, w) -> list of fused tokens bf16 [B*h*w, C]."""
        items = list(zip(self.levels, fx, fy))
        if os.environ.get("MMSAM_NECK_STREAMS", "1") == "0" or len(items) < 2 or debug is not None:
            return [self._level(lv, tx, ty, B, h, w, debug) for lv, (tx, h, w), (ty, _, _) in items]
        # The four pyramid levels are independent until the backbone consumes them, and the kernels of the small levels
        # (32^2 / 64^2 / 128^2 maps: a handful of CTAs each, plus the O(B*C^2) glue) leave most of the GPU idle: they
        # run on side streams next to the 256^2 level (fork / join on the current stream; inside a CUDA graph capture
        # this records parallel branches). Every tensor a side stream allocates is either freed on that stream or
        # returned and consumed after the join, and the next call forks again before reusing the side pools.
        main = torch.cuda.current_stream()
        if len(self._side) < len(items) - 1:
            self._side = [torch.cuda.Stream(device=self.dev) for _ in range(len(items) - 1)]
        outs = [None] * len(items)
        for i in range(1, len(items)):
            lv, (tx, h, w), (ty, _, _) = items[i]
            side = self._side[i - 1]
            side.wait_stream(main)
            with torch.cuda.stream(side):
                outs[i] = self._level(lv, tx, ty, B, h, w)
        lv, (tx, h, w), (ty, _, _) = items[0]
        outs[0] = self._level(lv, tx, ty, B, h, w)
        for side in self._side[: len(items) - 1]:
            main.wait_stream(side)
        return outs
