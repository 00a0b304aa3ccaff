"""Generate tests/golden/*.pt by running the UNMODIFIED reference (via tools/ref_shim.py) in the build
container. The reference tree does not travel to the GPU box; these small fixtures do.

    PYTHONDONTWRITEBYTECODE=1 python tools/make_golden.py

Fixtures:
  msda_known_answer.pt   ops/test.py's own vectors (seed 3) through ms_deform_attn_core_pytorch, fp64 + fp32
  msda_shapes.pt         larger / non-square / out-of-range sampling through ms_deform_attn_core_pytorch
  tiny_backbone.pt       full backbone forward of the reference class on the TINY config (fp32), all 4 outputs
  block_nonsquare.pt     reference Block (window + global, non-square 19x25 tokens, interpolated rel-pos)
  interaction_nonsquare.pt  reference InteractionBlock(extra_extractor=True) on a 320x448 image geometry
  vitb512_samples.pt     BASELINE config 1 (ViT-B 512^2): 16384 sampled outputs per feature map + norms
Weights are NOT stored: tests rebuild them deterministically (tests/common.py) and check a sha256.
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import ref_shim  # noqa: E402
import common  # noqa: E402
from oracle.perturb import synthetic_batch  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def msda():
    core = ref_shim.ref_msda_core()
    N, M, D, Lq, L, P = 1, 2, 2, 2, 2, 2
    shapes = torch.as_tensor([(6, 4), (3, 2)], dtype=torch.long)
    S = int(shapes.prod(1).sum())
    torch.manual_seed(3)
    rec = {}
    for name, dt in (("double", torch.float64), ("float", torch.float32)):
        value = torch.rand(N, S, M, D) * 0.01
        loc = torch.rand(N, Lq, M, L, P, 2)
        aw = torch.rand(N, Lq, M, L, P) + 1e-5
        aw /= aw.sum(-1, keepdim=True).sum(-2, keepdim=True)
        rec[name] = dict(value=value, loc=loc, aw=aw, out=core(value.to(dt), shapes, loc.to(dt), aw.to(dt)))
    rec["shapes"] = shapes
    torch.save(rec, os.path.join(OUT, "msda_known_answer.pt"))
    g = torch.Generator().manual_seed(7)
    cases = []
    for (N, M, D, Lq, shp, P) in ((2, 4, 8, 40, [(16, 12), (8, 6), (4, 3)], 4), (1, 3, 6, 17, [(7, 9)], 2),
                                 (1, 16, 32, 64, [(8, 8), (4, 4), (2, 2)], 4)):
        shapes = torch.as_tensor(shp, dtype=torch.long)
        S = int(shapes.prod(1).sum())
        value = torch.randn(N, S, M, D, generator=g, dtype=torch.float64)
        loc = torch.rand(N, Lq, M, len(shp), P, 2, generator=g, dtype=torch.float64) * 1.4 - 0.2
        aw = torch.rand(N, Lq, M, len(shp), P, generator=g, dtype=torch.float64)
        cases.append(dict(value=value.float(), shapes=shapes, loc=loc.float(), aw=aw.float(),
                          out=core(value, shapes, loc, aw).float()))
    torch.save(cases, os.path.join(OUT, "msda_shapes.pt"))


def backbone(cfg, hcfg, name, sample=None):
    seg, sd = common.build_segmentor(cfg, hcfg)
    bsd = {k[len("backbone."):]: v for k, v in sd.items() if k.startswith("backbone.")}
    net = ref_shim.build_backbone(cfg)
    net.load_state_dict(bsd, strict=True)
    x = synthetic_batch(1, cfg["img_size"])
    with torch.no_grad():
        outs, _ = net(x)
    rec = dict(digest=common.sd_digest(sd), norms=[o.norm().item() for o in outs])
    if sample is None:
        rec["outs"] = [o.clone() for o in outs]
    else:
        g = torch.Generator().manual_seed(11)
        idx = [torch.randint(0, o.numel(), (sample,), generator=g) for o in outs]
        rec["idx"] = idx
        rec["vals"] = [o.reshape(-1)[i].clone() for o, i in zip(outs, idx)]
    torch.save(rec, os.path.join(OUT, name))
    print(name, [tuple(o.shape) for o in outs])


def block_level():
    ref_shim.install()
    from mmseg_custom.models.backbones.base.image_encoder import Block
    from mmseg_custom.models.backbones.adapter_modules_multimodal_mix_mod_new_in_twin_convnext_new import \
        InteractionBlock, deform_inputs
    from functools import partial
    from oracle.perturb import perturb_state_dict
    import mmsam_b200  # noqa
    from mmsam_b200 import nn_modules as M
    from mmsam_b200.ops.modules import MSDeformAttn
    H, W, dim, nh = 19, 25, 128, 2
    g = torch.Generator().manual_seed(5)
    x = torch.randn(2, H * W, dim, generator=g)
    rec = dict(x=x, H=H, W=W)
    for name, ws in (("window", 14), ("global", 0)):
        torch.manual_seed(21)
        mine = M.Block(dim, nh, 4.0, True, True, ws, (64, 64))   # 127-row tables -> interpolated for global
        sd = perturb_state_dict(mine.state_dict(), seed=3)
        blk = Block(dim=dim, num_heads=nh, mlp_ratio=4.0, qkv_bias=True, norm_layer=partial(torch.nn.LayerNorm, eps=1e-6),
                    use_rel_pos=True, window_size=ws, input_size=(64, 64), with_cp=False).eval()
        blk.load_state_dict(sd, strict=True)
        with torch.no_grad():
            rec[name] = blk(x, H, W)
        rec[name + "_digest"] = common.sd_digest(sd)
    torch.save(rec, os.path.join(OUT, "block_nonsquare.pt"))
    # InteractionBlock on a 320 x 448 image geometry (tokens 20 x 28; H/32, W/32 must be integral for DWConv)
    Hi, Wi = 320, 448
    torch.manual_seed(22)
    mine = M.InteractionBlock(dim, 2, 4, True, 0.25, 0.5, 0.5, True, MSDeformAttn)
    sd = perturb_state_dict(mine.state_dict(), seed=4)
    ib = InteractionBlock(dim=dim, num_heads=2, n_points=4, init_values=0.5, drop_path=0.0,
                          norm_layer=partial(torch.nn.LayerNorm, eps=1e-6), with_cffn=True, cffn_ratio=0.25,
                          deform_ratio=0.5, extra_extractor=True, with_cp=False).eval()
    ib.load_state_dict(sd, strict=True)
    S3 = (Hi // 8) * (Wi // 8) + (Hi // 16) * (Wi // 16) + (Hi // 32) * (Wi // 32)
    xq = torch.randn(1, (Hi // 16) * (Wi // 16), dim, generator=g)
    c = torch.randn(1, S3, dim, generator=g)
    d1, d2 = deform_inputs(torch.zeros(1, 3, Hi, Wi))
    with torch.no_grad():
        xo, co = ib(xq, c, [], d1, d2, Hi // 16, Wi // 16)
    torch.save(dict(x=xq, c=c, Hi=Hi, Wi=Wi, xo=xo, co=co, digest=common.sd_digest(sd)),
               os.path.join(OUT, "interaction_nonsquare.pt"))


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    msda()
    block_level()
    backbone(common.TINY, common.TINY_HEAD, "tiny_backbone.pt")
    backbone(common.VITB, common.VITB_HEAD, "vitb512_samples.pt", sample=16384)
