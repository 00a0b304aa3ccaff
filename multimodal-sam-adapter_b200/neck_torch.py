"""RoadFormer2Neck fusion (adapter_modules_multimodal_mix_mod_new_in_twin_convnext_new.py:75-394) on the
GPU through ATen/cuDNN/cuBLAS library calls.

STATUS: this is the one stage of the path that is NOT yet hand-written sm_100a code (3 % of the
FLOPs, ~10 distinct small ops: grouped 1x1/3x3 convs, HW-long channel attentions, GroupNorm,
coordinate attention). It runs on the device (no CPU fallback) in fp32 for the ill-conditioned
HW-long softmax/LayerNorm statistics; DESIGN.md lists it under "library calls still on the path".
"""
import torch
import torch.nn.functional as F


def _p(t, dev):
    return t.detach().to(device=dev, dtype=torch.float32).contiguous()


class NeckTorch:
    def __init__(self, m, dev):
        self.levels = []
        for i in range(len(m.in_channels)):
            lv = {}
            for name, key in (("rgb", "global_feature_encoder_rgb"), ("sne", "global_feature_encoder_sne")):
                g = getattr(m, key)[i]
                lv["gfe_" + name] = dict(nw=_p(g.norm1.body.weight, dev), nb=_p(g.norm1.body.bias, dev),
                                         scale=_p(g.attn.scale, dev), scale2=_p(g.attn.scale2, dev),
                                         qkv1=_p(g.attn.qkv1.weight, dev), qkv2=_p(g.attn.qkv2.weight, dev),
                                         proj=_p(g.attn.proj.weight, dev), heads=g.attn.num_heads,
                                         groups=g.attn.qkv1.groups)
            for name, key in (("rgb", "local_feature_encoder_rgb"), ("sne", "local_feature_encoder_sne")):
                l = getattr(m, key)[i]
                lv["mb_" + name] = dict(w0=_p(l.bottleneckBlock[0].weight, dev), w2=_p(l.bottleneckBlock[2].weight, dev),
                                        w4=_p(l.bottleneckBlock[4].weight, dev), scale=_p(l.scale, dev))
            f = m.fuse_blocks[i]
            lv["gffm"] = dict(gx=_p(f.gammax.scale, dev), gy=_p(f.gammay.scale, dev), nw=_p(f.norm.weight, dev),
                              nb=_p(f.norm.bias, dev), eps=f.norm.eps)
            d = m.detail_feature_extractions[i]
            lv["mlp"] = dict(pin=_p(d.project_in.weight, dev), dw=_p(d.dwconv.weight, dev), pout=_p(d.project_out.weight, dev))
            e = m.enhance_blocks[i].conv_atten
            lv["ffrm"] = dict(w=_p(e.conv.weight, dev), gw=_p(e.gn.weight, dev), gb=_p(e.gn.bias, dev), groups=e.gn.num_groups,
                              eps=e.gn.eps)
            s = m.scale_layers[i]
            lv["s1"], lv["s2"] = _p(s.scale1, dev), _p(s.scale2, dev)
            ca = m.ca_blocks[i].coord_atten
            bn = ca.bn1
            bs = bn.weight.detach().float() / torch.sqrt(bn.running_var.detach().float() + bn.eps)
            bt = bn.bias.detach().float() - bn.running_mean.detach().float() * bs
            lv["ca"] = dict(w1=_p(ca.conv1.weight, dev), b1=_p(ca.conv1.bias, dev), bs=_p(bs, dev), bt=_p(bt, dev),
                            wh=_p(ca.conv_h.weight, dev), bh=_p(ca.conv_h.bias, dev), ww=_p(ca.conv_w.weight, dev),
                            bw=_p(ca.conv_w.bias, dev))
            self.levels.append(lv)

    @staticmethod
    def _gfe(x, p):
        b, c, h, w = x.shape
        t = x.flatten(2).transpose(1, 2)
        mu = t.mean(-1, keepdim=True)
        var = t.var(-1, keepdim=True, unbiased=False)
        n = ((t - mu) / torch.sqrt(var + 1e-5) * p["nw"] + p["nb"]).transpose(1, 2).reshape(b, c, h, w)
        qkv = F.conv2d(F.conv2d(n, p["qkv1"], None, groups=p["groups"]), p["qkv2"], None, padding=1, groups=p["groups"])
        q, k, v = qkv.chunk(3, dim=1)
        hd = p["heads"]
        q, k, v = (z.reshape(b, hd, c // hd, h * w) for z in (q, k, v))
        q, k = F.normalize(q, dim=-1), F.normalize(k, dim=-1)
        att = ((q @ k.transpose(-2, -1)) * p["scale"]).softmax(-1)
        o = F.conv2d((att @ v).reshape(b, c, h, w), p["proj"])
        return x + n + o * p["scale2"]

    @staticmethod
    def _mb(x, p):
        y = F.relu6(F.conv2d(x, p["w0"]))
        y = F.relu6(F.conv2d(y, p["w2"], None, padding=1, groups=y.shape[1]))
        return F.conv2d(y, p["w4"]) * p["scale"] + x

    @staticmethod
    def _gffm(x, p):
        b, c2, h, w = x.shape
        c = c2 // 2
        fx, fy = x[:, :c].reshape(b, c, -1), x[:, c:].reshape(b, c, -1)
        ax = F.softmax(torch.bmm(fx, fy.transpose(1, 2)), -1)
        ay = F.softmax(torch.bmm(fy, fx.transpose(1, 2)), -1)
        ox = torch.bmm(ax, fy) * p["gx"] + fx
        oy = torch.bmm(ay, fx) * p["gy"] + fy
        o = F.layer_norm(torch.cat((ox, oy), 1), (h * w,), p["nw"], p["nb"], p["eps"])
        return o.view(b, c2, h, w)

    @staticmethod
    def _mlp(x, p):
        y = F.conv2d(x, p["pin"])
        y = F.conv2d(y, p["dw"], None, padding=1, groups=y.shape[1] // 2)
        a, g = y.chunk(2, 1)
        return F.conv2d(F.gelu(a) * g, p["pout"])

    @staticmethod
    def _ffrm(x, p):
        a = F.conv2d(F.adaptive_avg_pool2d(x, 1), p["w"])
        a = F.relu(F.group_norm(a, p["groups"], p["gw"], p["gb"], p["eps"]))
        return x + x * torch.sigmoid(a)

    @staticmethod
    def _ca(x, p):
        n, c, h, w = x.shape
        y = torch.cat((x.mean(3, keepdim=True), x.mean(2, keepdim=True).permute(0, 1, 3, 2)), 2)
        y = F.conv2d(y, p["w1"], p["b1"]) * p["bs"].view(1, -1, 1, 1) + p["bt"].view(1, -1, 1, 1)
        y = y * F.relu6(y + 3) / 6
        yh, yw = torch.split(y, [h, w], 2)
        ah = torch.sigmoid(F.conv2d(yh, p["wh"], p["bh"]))
        aw = torch.sigmoid(F.conv2d(yw.permute(0, 1, 3, 2), p["ww"], p["bw"]))
        return x + x * aw * ah

    @torch.no_grad()
    def __call__(self, fx, fy, B):
        """fx / fy: per level (tokens bf16 [B*h*w, Ci], h, w) -> list of fused tokens bf16 [B*h*w, 2*Ci]."""
        # TF32 tensor-core library kernels for the convs / bmm (10-bit mantissa inputs, fp32 accumulate: finer
        # than the bf16 used everywhere else on the path); statistics, softmax and norms stay fp32.
        prev = torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32
        torch.backends.cudnn.allow_tf32 = True
        torch.backends.cuda.matmul.allow_tf32 = True
        try:
            outs = []
            for lv, (tx, h, w), (ty, _, _) in zip(self.levels, fx, fy):
                ci = tx.shape[1]
                rgb = tx.view(B, h, w, ci).permute(0, 3, 1, 2).float()
                aux = ty.view(B, h, w, ci).permute(0, 3, 1, 2).float()
                g = torch.cat((self._gfe(rgb, lv["gfe_rgb"]), self._gfe(aux, lv["gfe_sne"])), 1)
                l = torch.cat((self._mb(rgb, lv["mb_rgb"]), self._mb(aux, lv["mb_sne"])), 1)
                g = self._ffrm(self._gffm(g, lv["gffm"]), lv["ffrm"])
                l = self._mlp(l, lv["mlp"])
                f = self._ca(g * lv["s1"] + l * lv["s2"], lv["ca"])
                outs.append(f.permute(0, 2, 3, 1).reshape(B * h * w, 2 * ci).to(torch.bfloat16).contiguous())
            return outs
        finally:
            torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = prev
