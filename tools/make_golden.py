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
  segmentor_tiny.pt      reference EncoderDecoder + SegformerHead (unmodified; base classes stubbed, tools/ref_shim.py) on
                         the TINY config: head logits and simple_test labels for whole_dim (dim == / != input), whole_dim_cut
                         (rescale on / off), whole (ori_shape != input), slide (+ rescale), horizontal / vertical flip
  head_vitl.pt           reference SegformerHead.forward at the ViT-L head size (4 x 1024 -> 512 -> 25) on random features
  blocks_68x120.pt       BASELINE config 5b: reference Block (window + global) and InteractionBlock(extra_extractor=True)
                         at 68 x 120 tokens (1088 x 1920 geometry), dim 128: sampled outputs + norms
  fmb800_samples.pt      BASELINE config 4: reference ...NEWwithcp ViT-L, 800 x 800, 14 classes, whole_dim_cut: sampled
                         features / logits, full label map [600, 800]
  vitl1024_samples.pt    BASELINE config 2: reference ViT-L 1024 x 1024, 25 classes: sampled features / logits, full labels
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


def _sample(t, n, seed):
    g = torch.Generator().manual_seed(seed)
    idx = torch.randint(0, t.numel(), (min(n, t.numel()),), generator=g)
    return idx, t.reshape(-1)[idx].clone()


def segmentor_tiny():
    """The reference's own EncoderDecoder.simple_test / SegformerHead.forward on the TINY config, every inference mode."""
    import numpy as np
    seg, sd = common.build_segmentor(common.TINY, common.TINY_HEAD)
    net = ref_shim.build_segmentor(common.TINY, common.TINY_HEAD, dict(mode="whole"))
    net.load_state_dict(sd, strict=True)
    Dict = sys.modules["addict"].Dict
    x = synthetic_batch(2, 128, seed=5)
    rec = dict(digest=common.sd_digest(sd), cases={})
    with torch.no_grad():
        feats, _ = net.backbone(x)
        rec["head_logits"] = net.decode_head(feats).clone()                 # [2, 25, 32, 32]
        rec["logits_img_idx"], rec["logits_img_vals"] = _sample(net.encode_decode(x, None), 65536, 51)   # of [2, 25, 128, 128]

        def run(name, test_cfg, img, rescale, ori_shape=None, flip=False, direction="horizontal"):
            net.test_cfg = Dict(test_cfg)
            meta = [dict(ori_shape=tuple(ori_shape or img.shape[2:]) + (3,), flip=flip, flip_direction=direction)] * img.shape[0]
            out = net.simple_test(img, meta, rescale)
            rec["cases"][name] = dict(test_cfg=dict(test_cfg), rescale=rescale, ori_shape=ori_shape, flip=flip,
                                      direction=direction, labels=torch.as_tensor(np.stack(out)).to(torch.uint8))
        run("whole_dim", dict(mode="whole_dim", dim=(128, 128)), x, True)
        run("whole_dim_resized", dict(mode="whole_dim", dim=(96, 160)), x, True)
        run("whole_dim_cut", dict(mode="whole_dim_cut", dim=(96, 128), cut_dim=(128, 96)), x, False)
        run("whole_dim_cut_rescaled", dict(mode="whole_dim_cut", dim=(160, 144), cut_dim=(120, 100)), x, True)
        run("whole", dict(mode="whole"), x, True, ori_shape=(100, 150))
        run("whole_norescale", dict(mode="whole"), x, False)
        run("whole_flip_h", dict(mode="whole"), x, True, flip=True)
        run("whole_dim_flip_v", dict(mode="whole_dim", dim=(128, 128)), x, True, flip=True, direction="vertical")
        frame = torch.cat([synthetic_batch(1, 128, seed=21), synthetic_batch(1, 128, seed=22)], 3)[:, :, :, :208]
        rec["frame_seeds"] = (21, 22)
        run("slide", dict(mode="slide", crop_size=(128, 128), stride=(64, 80)), frame, False)
        run("slide_rescaled", dict(mode="slide", crop_size=(128, 128), stride=(64, 80)), frame, True, ori_shape=(96, 160))
        net.test_cfg = Dict(mode="slide", crop_size=(128, 128), stride=(64, 80))
        rec["slide_idx"], rec["slide_vals"] = _sample(net.slide_inference(frame, [dict(ori_shape=(128, 208, 3))], False), 65536, 52)
    torch.save(rec, os.path.join(OUT, "segmentor_tiny.pt"))
    print("segmentor_tiny.pt", {k: tuple(v["labels"].shape) for k, v in rec["cases"].items()})


def head_vitl():
    import bench
    from oracle.perturb import perturb_state_dict
    import mmsam_b200  # noqa
    from mmsam_b200.backbone import SegformerHead
    torch.manual_seed(31)
    mine = SegformerHead(**bench.VITL_HEAD)
    sd = perturb_state_dict(mine.state_dict(), seed=6)
    head = ref_shim.build_head(bench.VITL_HEAD)
    head.load_state_dict(sd, strict=True)
    g = torch.Generator().manual_seed(32)
    feats = [torch.randn(2, 1024, 64 >> i, 64 >> i, generator=g) for i in range(4)]
    with torch.no_grad():
        out = head(feats)
    torch.save(dict(digest=common.sd_digest(sd), seed=32, out=out.clone()), os.path.join(OUT, "head_vitl.pt"))
    print("head_vitl.pt", tuple(out.shape))


def blocks_68x120():
    """BASELINE config 5b: ViT blocks + InteractionBlock at 68 x 120 = 8160 tokens."""
    ref_shim.install()
    from mmseg_custom.models.backbones.base.image_encoder import Block
    from mmseg_custom.models.backbones.adapter_modules_multimodal_mix_mod_new_in_twin_convnext_new import \
        InteractionBlock, deform_inputs
    from functools import partial
    from oracle.perturb import perturb_state_dict
    import mmsam_b200  # noqa
    from mmsam_b200 import nn_modules as M
    from mmsam_b200.ops.modules import MSDeformAttn
    H, W, dim, nh = 68, 120, 128, 2
    g = torch.Generator().manual_seed(41)
    x = torch.randn(1, H * W, dim, generator=g)
    rec = dict(H=H, W=W, dim=dim, nh=nh, x_seed=41)
    for name, ws in (("window", 14), ("global", 0)):
        torch.manual_seed(42)
        mine = M.Block(dim, nh, 4.0, True, True, ws, (64, 64))
        sd = perturb_state_dict(mine.state_dict(), seed=7)
        blk = Block(dim=dim, num_heads=nh, mlp_ratio=4.0, qkv_bias=True, norm_layer=partial(torch.nn.LayerNorm, eps=1e-6),
                    use_rel_pos=True, window_size=ws, input_size=(64, 64), with_cp=False).eval()
        blk.load_state_dict(sd, strict=True)
        with torch.no_grad():
            out = blk(x, H, W)
        idx, vals = _sample(out, 65536, 43)
        rec[name] = dict(idx=idx, vals=vals, norm=out.norm().item(), digest=common.sd_digest(sd))
    Hi, Wi = 16 * H, 16 * W
    torch.manual_seed(44)
    mine = M.InteractionBlock(dim, nh, 4, True, 0.25, 0.5, 0.5, True, MSDeformAttn)
    sd = perturb_state_dict(mine.state_dict(), seed=8)
    ib = InteractionBlock(dim=dim, num_heads=nh, n_points=4, init_values=0.5, drop_path=0.0,
                          norm_layer=partial(torch.nn.LayerNorm, eps=1e-6), with_cffn=True, cffn_ratio=0.25,
                          deform_ratio=0.5, extra_extractor=True, with_cp=False).eval()
    ib.load_state_dict(sd, strict=True)
    S3 = (Hi // 8) * (Wi // 8) + (Hi // 16) * (Wi // 16) + (Hi // 32) * (Wi // 32)
    g = torch.Generator().manual_seed(45)
    xq = torch.randn(1, H * W, dim, generator=g)
    c = torch.randn(1, S3, dim, generator=g)
    d1, d2 = deform_inputs(torch.zeros(1, 3, Hi, Wi))
    with torch.no_grad():
        xo, co = ib(xq, c, [], d1, d2, H, W)
    ix, vx = _sample(xo, 65536, 46)
    ic, vc = _sample(co, 65536, 47)
    rec["interaction"] = dict(seed=45, Hi=Hi, Wi=Wi, S3=S3, idx_x=ix, vals_x=vx, norm_x=xo.norm().item(), idx_c=ic, vals_c=vc,
                              norm_c=co.norm().item(), digest=common.sd_digest(sd))
    torch.save(rec, os.path.join(OUT, "blocks_68x120.pt"))
    print("blocks_68x120.pt")


def full_segmentor(cfg, hcfg, name, size, kind, test_cfg, rescale, withcp, btype, zero_rows_from=None, nsample=16384):
    """Full-size reference run (backbone + head + EncoderDecoder.simple_test), 1 image: sampled feature maps and logits,
    the complete label map."""
    import numpy as np
    seg, sd = common.build_segmentor(cfg, hcfg, btype=btype)
    del seg
    net = ref_shim.build_segmentor(cfg, hcfg, test_cfg, withcp=withcp)
    net.load_state_dict(sd, strict=True)
    x = synthetic_batch(1, size, kind=kind)
    if zero_rows_from is not None:
        x[:, :, zero_rows_from:] = 0
    with torch.no_grad():
        feats, _ = net.backbone(x)
        hl = net.decode_head(feats)
        labels = net.simple_test(x, [dict(ori_shape=(size, size, 3), flip=False)], rescale)
    rec = dict(digest=common.sd_digest(sd), norms=[f.norm().item() for f in feats], head_norm=hl.norm().item(),
               labels=torch.as_tensor(np.stack(labels)).to(torch.uint8), test_cfg=dict(test_cfg), rescale=rescale)
    smp = [_sample(f, nsample, 11 + i) for i, f in enumerate(feats)]
    rec["idx"], rec["vals"] = [a for a, _ in smp], [b for _, b in smp]
    rec["head_idx"], rec["head_vals"] = _sample(hl, 4 * nsample, 17)
    torch.save(rec, os.path.join(OUT, name))
    print(name, [tuple(f.shape) for f in feats], tuple(rec["labels"].shape))


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    if len(sys.argv) > 1:                       # regenerate selected fixtures only: python tools/make_golden.py segmentor_tiny
        for fn in sys.argv[1:]:
            globals()[fn]()
        sys.exit(0)
    msda()
    block_level()
    backbone(common.TINY, common.TINY_HEAD, "tiny_backbone.pt")
    backbone(common.VITB, common.VITB_HEAD, "vitb512_samples.pt", sample=16384)
    segmentor_tiny()
    head_vitl()
    blocks_68x120()
    import bench
    full_segmentor(common.FMB, common.FMB_HEAD, "fmb800_samples.pt", 800, "thermal", common.FMB_TEST_CFG, False, True,
                   "SAMAdapterbimodalMixModNewInTwinConvNEWwithcp", zero_rows_from=600)
    full_segmentor(bench.VITL, bench.VITL_HEAD, "vitl1024_samples.pt", 1024, "lidar",
                   dict(mode="whole_dim", rescale=True, dim=(1024, 1024)), True, False, "SAMAdapterbimodalMixModNewInTwinConvNEW")
