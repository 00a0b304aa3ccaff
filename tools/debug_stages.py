"""Stage-by-stage comparison of the CUDA path with the CPU oracle (GPU box)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from common import TINY, TINY_HEAD, VITB, VITB_HEAD, build_segmentor, rel_l2  # noqa
from oracle import model as om  # noqa
from oracle.perturb import synthetic_batch  # noqa

which = sys.argv[1] if len(sys.argv) > 1 else "tiny"
if which == "vitl":
    import bench
    cfg, hcfg = bench.VITL, bench.VITL_HEAD
    torch.set_num_threads(os.cpu_count() or 1)
else:
    cfg, hcfg = (TINY, TINY_HEAD) if which == "tiny" else (VITB, VITB_HEAD)
seg, sd = build_segmentor(cfg, hcfg)
x = synthetic_batch(1, cfg["img_size"])
st = {}
with torch.no_grad():
    want = om.backbone_forward(sd, cfg, x, prefix="backbone.", stages=st)
seg = seg.cuda()
dbg = {}
eng = seg.backbone.engine(seg.decode_head)
got = eng.backbone_nhwc(x.cuda(), debug=dbg)


def tok(t):  # NCHW -> [HW, C]
    return t[0].flatten(1).t()


for i in range(4):
    tw = st["twin"][i]
    ci = tw.shape[1] // 2
    print(f"twin level {i}: rgb {rel_l2(dbg['fx'][i][0].float().cpu(), tok(tw[:, :ci])):.3e} aux {rel_l2(dbg['fy'][i][0].float().cpu(), tok(tw[:, ci:])):.3e}")
for i in range(4):
    nd, no = dbg["neck"][i], st["neck"][i]
    print(f"neck level {i}: g {rel_l2(nd['g'].float().cpu(), tok(no['g'])):.3e} l {rel_l2(nd['l'].float().cpu(), tok(no['l'])):.3e} "
          f"lo {rel_l2(nd['lo'].float().cpu(), tok(no['lo'])):.3e} f {rel_l2(nd['f'].float().cpu(), tok(no['f'])):.3e} "
          f"fused {rel_l2(dbg['fused'][i].float().cpu(), tok(st['fused'][i])):.3e}")
print("c1", rel_l2(dbg["c1"].float().cpu(), st["c1"][0]))
print("c_0", rel_l2(dbg["c_0"].float().cpu(), st["c_0"][0]))
print("x_0", rel_l2(dbg["x_0"].float().cpu(), st["x_0"][0]))
for i in range(4):
    print(f"x{i}", rel_l2(dbg[f"x{i}"].float().cpu(), st[f"x{i}"][0]), f"c_{i}", rel_l2(dbg[f"c_{i}"].float().cpu(), st[f"c_{i}"][0]))
for i, (g, w) in enumerate(zip(got, want)):
    print(f"f{i + 1}", rel_l2(g.permute(0, 3, 1, 2).float().cpu(), w))

# ---- inside GFFM / FFRM (adapter_modules_...new.py:242-267, 148-162): where does `f` pick up its error? ----
import torch.nn.functional as F  # noqa: E402
s_ = om.SD(sd, "backbone.spm.smart_fusion.")
for i in range(4):
    nd, no = dbg["neck"][i], st["neck"][i]
    g_ref = no["g"].double()                                     # [1, C, h, w] oracle input of GFFM
    b, c2, h, w = g_ref.shape
    c = c2 // 2
    fs = s_.sub(f"fuse_blocks.{i}")

    def gffm_pre(gx):
        fx, fy = gx[:, :c].reshape(b, c, -1), gx[:, c:].reshape(b, c, -1)
        E = torch.bmm(fx, fy.transpose(1, 2))
        ax, ay = F.softmax(E, -1), F.softmax(E.transpose(1, 2), -1)
        return torch.cat((torch.bmm(ax, fy) * fs("gammax.scale").double() + fx, torch.bmm(ay, fx) * fs("gammay.scale").double() + fy), 1), E
    o_ref, E_ref = gffm_pre(g_ref)
    g_gpu = nd["g"].double().cpu().t().reshape(1, c2, h, w)
    o_from_gpu_g, E_gpu_g = gffm_pre(g_gpu)                      # exact arithmetic on the GPU's (bf16) g: pure input sensitivity
    o_gpu = nd["o"].double().cpu().t().reshape(1, c2, h * w)
    mu_ref, var_ref = o_ref.mean(-1), o_ref.var(-1, unbiased=False)
    print(f"gffm level {i}: E max|.| {E_ref.abs().max():.3g} softmax row max (mean) {F.softmax(E_ref, -1).max(-1).values.mean():.3f} | "
          f"o: exact-on-GPU-g {rel_l2(o_from_gpu_g, o_ref):.3e} GPU {rel_l2(o_gpu, o_ref):.3e} | "
          f"mu {rel_l2(nd['mu'].double().cpu(), mu_ref):.3e} rstd {rel_l2(nd['rstd'].double().cpu(), 1 / torch.sqrt(var_ref + 1e-5)):.3e} | "
          f"o - mu (centred) GPU {rel_l2(o_gpu - nd['mu'].double().cpu()[..., None], o_ref - mu_ref[..., None]):.3e}")
