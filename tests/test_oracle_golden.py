"""CPU tests: the oracle against the golden vectors produced by the UNMODIFIED reference
(tools/make_golden.py), the C-ABI surface, and host-side logic. No GPU needed."""
import os
import re
import sys

import pytest
import torch

from common import TINY, TINY_HEAD, VITB, VITB_HEAD, build_segmentor, rel_l2, sd_digest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")


def test_msda_oracle_known_answer_vectors():
    """segmentation/ops/test.py:16-75 — the reference's only known-answer test."""
    from oracle.msda import ms_deform_attn_core
    rec = torch.load(os.path.join(GOLD, "msda_known_answer.pt"))
    d = rec["double"]
    out = ms_deform_attn_core(d["value"].double(), rec["shapes"], d["loc"].double(), d["aw"].double())
    assert torch.allclose(out, d["out"])                         # ops/test.py:44 criterion
    f = rec["float"]
    out = ms_deform_attn_core(f["value"], rec["shapes"], f["loc"], f["aw"])
    assert torch.allclose(out, f["out"], rtol=1e-2, atol=1e-3)  # ops/test.py:68 criterion
    assert (out - f["out"]).abs().max() < 1e-8


def test_msda_oracle_shapes():
    from oracle.msda import ms_deform_attn_core
    for c in torch.load(os.path.join(GOLD, "msda_shapes.pt")):
        out = ms_deform_attn_core(c["value"].double(), c["shapes"], c["loc"].double(), c["aw"].double())
        assert (out.float() - c["out"]).abs().max() < 1e-5


def test_oracle_tiny_backbone_matches_reference():
    from oracle import model as om
    from oracle.perturb import synthetic_batch
    rec = torch.load(os.path.join(GOLD, "tiny_backbone.pt"))
    _, sd = build_segmentor(TINY, TINY_HEAD)
    assert sd_digest(sd) == rec["digest"], "deterministic weight construction drifted from the golden run"
    with torch.no_grad():
        outs = om.backbone_forward(sd, TINY, synthetic_batch(1, TINY["img_size"]), prefix="backbone.")
    for o, r in zip(outs, rec["outs"]):
        assert rel_l2(o, r) < 1e-5


def test_oracle_block_nonsquare_matches_reference():
    """Window (padded 19x25 -> 28x28) and global (127-row table interpolated to 37 / 49 rows) blocks."""
    from oracle import model as om
    from oracle.perturb import perturb_state_dict
    import mmsam_b200  # noqa
    from mmsam_b200 import nn_modules as M
    rec = torch.load(os.path.join(GOLD, "block_nonsquare.pt"))
    for name, ws in (("window", 14), ("global", 0)):
        torch.manual_seed(21)
        sd = perturb_state_dict(M.Block(128, 2, 4.0, True, True, ws, (64, 64)).state_dict(), seed=3)
        assert sd_digest(sd) == rec[name + "_digest"]
        with torch.no_grad():
            out = om.vit_block(rec["x"], om.SD(sd), rec["H"], rec["W"], ws, 2)
        assert rel_l2(out, rec[name]) < 1e-5


def test_oracle_interaction_nonsquare_matches_reference():
    from oracle import model as om
    from oracle.perturb import perturb_state_dict
    import mmsam_b200  # noqa
    from mmsam_b200 import nn_modules as M
    from mmsam_b200.ops.modules import MSDeformAttn
    rec = torch.load(os.path.join(GOLD, "interaction_nonsquare.pt"))
    torch.manual_seed(22)
    sd = perturb_state_dict(M.InteractionBlock(128, 2, 4, True, 0.25, 0.5, 0.5, True, MSDeformAttn).state_dict(), seed=4)
    assert sd_digest(sd) == rec["digest"]
    s = om.SD(sd)
    Hi, Wi = rec["Hi"], rec["Wi"]
    (ref1, sh1), (ref2, sh2) = om.deform_inputs(Hi, Wi, torch.float32)
    with torch.no_grad():
        x = om.injector(rec["x"], rec["c"], ref1, sh1, s.sub("injector"), 2, 4)
        c = om.extractor(rec["c"], x, ref2, sh2, s.sub("extractor"), 2, 4, Hi // 16, Wi // 16)
        for j in range(2):
            c = om.extractor(c, x, ref2, sh2, s.sub(f"extra_extractors.{j}"), 2, 4, Hi // 16, Wi // 16)
    assert rel_l2(x, rec["xo"]) < 1e-5 and rel_l2(c, rec["co"]) < 1e-5


@pytest.mark.timeout(600)
def test_oracle_vitb512_matches_reference_samples():
    """BASELINE config 1 (ViT-B MM-adapter, 512x512, fp32 CPU): sampled outputs of the reference run."""
    from oracle import model as om
    from oracle.perturb import synthetic_batch
    rec = torch.load(os.path.join(GOLD, "vitb512_samples.pt"))
    _, sd = build_segmentor(VITB, VITB_HEAD)
    assert sd_digest(sd) == rec["digest"]
    with torch.no_grad():
        outs = om.backbone_forward(sd, VITB, synthetic_batch(1, 512), prefix="backbone.")
    for o, idx, vals, nrm in zip(outs, rec["idx"], rec["vals"], rec["norms"]):
        assert rel_l2(o.reshape(-1)[idx], vals) < 1e-4
        assert abs(o.norm().item() - nrm) / nrm < 1e-4


def test_state_dict_schema_matches_reference():
    """Key names + shapes of the ViT-L DELIVER backbone, dumped from the reference module."""
    import json
    import mmsam_b200  # noqa
    from mmsam_b200 import backbone  # noqa
    from mmsam_b200.registry import BACKBONES
    want = json.load(open(os.path.join(GOLD, "state_dict_schema_vitl_deliver.json")))
    cfg = dict(type="SAMAdapterbimodalMixModNewInTwinConvNEW", img_size=1024, modalities_name=["rgb", "lidar"],
               modalities_ch=[3, 3], init_values=1e-6, gamma_init_values=1e-6, patch_size=16, embed_dim=1024, depth=24,
               num_heads=16, mlp_ratio=4, drop_path_rate=0.3, drop_multimodal_path=0, conv_inplane=48, n_points=4,
               deform_num_heads=16, cffn_ratio=0.25, deform_ratio=0.5, with_cp=True,
               interaction_indexes=[[0, 5], [6, 11], [12, 17], [18, 23]], global_attn_indexes=[5, 11, 17, 23],
               window_size=14, arch="small", checkpoint="x")
    with torch.device("meta"):
        net = BACKBONES.build(cfg)
    got = {k: list(v.shape) for k, v in net.state_dict().items()}
    assert got == want


@pytest.mark.skipif(not os.path.isdir("/root/reference/segmentation/configs"), reason="reference tree not present")
def test_reference_configs_build_unchanged():
    """All 10 shipped configs load through the Config shim and name registered classes."""
    import glob
    import mmsam_b200  # noqa
    from mmsam_b200 import backbone  # noqa
    from mmsam_b200.registry import BACKBONES, HEADS, SEGMENTORS, load_config
    files = sorted(glob.glob("/root/reference/segmentation/configs/*/Segformer_MMSAM*.py"))
    assert len(files) == 10
    for f in files:
        cfg = load_config(f)
        assert SEGMENTORS.get(cfg.model.type) is not None
        assert BACKBONES.get(cfg.model.backbone.type) is not None
        assert HEADS.get(cfg.model.decode_head.type) is not None
        with torch.device("meta"):
            seg = SEGMENTORS.build({k: v for k, v in cfg.model.items() if k != "train_cfg"} | {"pretrained": None})
        assert seg.num_classes == cfg.model.decode_head.num_classes


def test_c_abi_exports_every_declared_symbol():
    import ctypes
    import mmsam_b200  # noqa
    from mmsam_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "mmsam_b200.h")).read()
    declared = set(re.findall(r"^int\s+(mmsam_\w+)\s*\(", hdr, flags=re.M))
    assert declared and declared == set(_lib.SIGNATURES), (declared ^ set(_lib.SIGNATURES))
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), name
    assert _lib.load().mmsam_arch() == 100
    # the ctypes stub must agree with every prototype: parameter count and the C type class of each parameter
    kind = {ctypes.c_void_p: "ptr", ctypes.c_int: "int", ctypes.c_longlong: "ll", ctypes.c_float: "float",
            ctypes.c_double: "double"}
    for name, params in re.findall(r"^int\s+(mmsam_\w+)\s*\(([^;]*?)\)\s*;", hdr, flags=re.M | re.S):
        plist = [q.strip() for q in params.replace("\n", " ").split(",")]
        if plist == ["void"]:
            plist = []
        want = []
        for q in plist:
            if "*" in q:
                want.append("ptr")
            elif re.match(r"^(const\s+)?long long\b", q):
                want.append("ll")
            elif re.match(r"^(const\s+)?float\b", q):
                want.append("float")
            elif re.match(r"^(const\s+)?double\b", q):
                want.append("double")
            else:
                assert re.match(r"^(const\s+)?int\b", q), (name, q)
                want.append("int")
        got = [kind[t] for t in _lib.SIGNATURES[name]]
        assert got == want, (name, got, want)


def test_no_cpu_fallback():
    import mmsam_b200  # noqa
    from mmsam_b200 import kernels
    from mmsam_b200._lib import MMSamError
    with pytest.raises(MMSamError):
        kernels.gemm(torch.zeros(8, 8, dtype=torch.bfloat16), torch.zeros(8, 8, dtype=torch.bfloat16))
    with pytest.raises(MMSamError):
        kernels.layernorm(torch.zeros(8, 8, dtype=torch.bfloat16), torch.ones(8), torch.zeros(8), 1e-6)


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "multimodal-sam-adapter_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith(".py"):
                src = open(os.path.join(dp, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", src, flags=re.M), f


def test_window_maps_roundtrip():
    import mmsam_b200  # noqa
    from mmsam_b200.engine import window_maps
    B, H, W, ws = 2, 19, 25, 14
    m = window_maps(B, H, W, ws, 8, "cpu")
    x = torch.arange(B * H * W, dtype=torch.float32)
    y = torch.zeros(m["win_rows"])
    y[m["win_fwd"].long()] = x
    ref = torch.nn.functional.pad(x.view(B, H, W), (0, 28 - W, 0, 28 - H))
    ref = ref.view(B, 2, ws, 2, ws).permute(0, 1, 3, 2, 4).reshape(-1)
    assert torch.equal(y, ref)
    inv = m["win_inv"].long()
    back = torch.zeros(B * H * W)
    back[inv[inv >= 0]] = y[inv >= 0]
    assert torch.equal(back, x)


def test_msdeformattn_module_contract():
    import mmsam_b200  # noqa
    from mmsam_b200.ops.modules import MSDeformAttn
    m = MSDeformAttn(d_model=64, n_levels=3, n_heads=4, n_points=4, ratio=0.5)
    assert m.im2col_step == 64 and set(dict(m.named_children())) == {"sampling_offsets", "attention_weights", "value_proj", "output_proj"}
    assert m.sampling_offsets.weight.abs().sum() == 0 and m.attention_weights.bias.abs().sum() == 0
    assert m.value_proj.weight.shape == (32, 64) and m.output_proj.weight.shape == (64, 32)
    with pytest.raises(ValueError):
        MSDeformAttn(d_model=65, n_heads=4)


def test_checkpoint_ingestion_rules():
    """mmcv-style unwrap / module. strip, the SAM image-encoder conversion
    (tools/SAM_checkpoint_convert.py:15-33) and the twin-ConvNeXt key surgery (base/twin_convnext.py:399-443)."""
    import mmsam_b200  # noqa: F401
    from mmsam_b200 import checkpoint as ck
    from mmsam_b200 import nn_modules as M
    torch.manual_seed(0)
    # (a) SAM: image_encoder.* minus neck.* -> ImageEncoderViT keys
    sam = {"image_encoder.pos_embed": torch.zeros(1), "image_encoder.blocks.0.norm1.weight": torch.ones(2),
           "image_encoder.neck.0.weight": torch.ones(3), "prompt_encoder.x": torch.ones(1), "mask_decoder.y": torch.ones(1)}
    assert list(ck.convert_sam_image_encoder(sam)) == ["pos_embed", "blocks.0.norm1.weight"]
    # (b) twin ConvNeXt from a single-tower checkpoint (wrapped + 'backbone.' prefix)
    twin = M.TwinConvNeXt(arch=dict(depths=[1, 1, 1, 1], channels=[8, 16, 32, 64]))
    single = {}
    for k, v in twin.state_dict().items():
        if "_x" in k.split(".")[0]:
            single["backbone." + k.replace("_x", "", 1)] = torch.randn_like(v)
    single["backbone.norm0.weight"] = torch.randn(8)          # per-stage output norms: norm{i} -> norm{i}_x ... see below
    left = ck.load_twin_convnext(twin, {"state_dict": single})
    sd = twin.state_dict()
    for k, v in single.items():
        k = k[len("backbone."):]
        dot = k.find(".")
        kx, ky = k[:dot] + "_x" + k[dot:], k[:dot] + "_y" + k[dot:]
        if kx in sd:
            assert torch.equal(sd[kx], v) and torch.equal(sd[ky], v)   # both towers got the same tensor
        else:
            assert kx in left                                           # e.g. norm0_x: the reference's own naming quirk
    # (c) mmcv-style load of a DataParallel-wrapped segmentor checkpoint
    lin = torch.nn.Linear(3, 2)
    wrapped = {"state_dict": {"module." + k: v + 1 for k, v in lin.state_dict().items()}, "meta": {}}
    missing, unexpected = ck.load_checkpoint(lin, wrapped)
    assert missing == [] and unexpected == []
    assert torch.equal(lin.weight.data, wrapped["state_dict"]["module.weight"])


# ---------------------------------------------------------------------------------------------------------------------
# Head / segmentor pin and the full-size configs: goldens produced by the reference's own
# SegformerHead.forward / EncoderDecoder.simple_test (tools/make_golden.py, tools/ref_shim.py)
# ---------------------------------------------------------------------------------------------------------------------
def _frame():
    from oracle.perturb import synthetic_batch
    return torch.cat([synthetic_batch(1, 128, seed=21), synthetic_batch(1, 128, seed=22)], 3)[:, :, :, :208]


def test_oracle_head_and_inference_modes_match_reference_segmentor():
    """decode_heads/segformer_head.py:48-66 + segmentors/encoder_decoder.py:191-234, 310-414, 417-508 on the TINY config."""
    from oracle import model as om
    from oracle.perturb import synthetic_batch
    rec = torch.load(os.path.join(GOLD, "segmentor_tiny.pt"))
    _, sd = build_segmentor(TINY, TINY_HEAD)
    assert sd_digest(sd) == rec["digest"]
    x = synthetic_batch(2, 128, seed=5)
    with torch.no_grad():
        feats = om.backbone_forward(sd, TINY, x, prefix="backbone.")
        hl = om.segformer_head(sd, feats)
        assert rel_l2(hl, rec["head_logits"]) < 1e-5
        assert rel_l2(om.encode_decode(sd, TINY, x).reshape(-1)[rec["logits_img_idx"]], rec["logits_img_vals"]) < 1e-5
        tc = dict(mode="slide", crop_size=(128, 128), stride=(64, 80))
        assert rel_l2(om.slide_inference(sd, TINY, _frame(), tc).reshape(-1)[rec["slide_idx"]], rec["slide_vals"]) < 1e-5
        for name, c in rec["cases"].items():
            img = _frame() if name.startswith("slide") else x
            got = om.simple_test(sd, TINY, img, c["test_cfg"], c["rescale"], c["ori_shape"], c["flip"], c["direction"])
            assert got.shape == c["labels"].shape, name
            agree = (got == c["labels"].long()).float().mean().item()
            assert agree >= 0.9995, (name, agree)      # fp32 summation-order ties only


def test_oracle_head_vitl_matches_reference_head():
    """The reference SegformerHead.forward at the ViT-L head size (4 x 1024 -> 512 -> fusion 2048 -> 512 -> 25)."""
    import bench
    from oracle import model as om
    from oracle.perturb import perturb_state_dict
    import mmsam_b200  # noqa
    from mmsam_b200.backbone import SegformerHead
    rec = torch.load(os.path.join(GOLD, "head_vitl.pt"))
    torch.manual_seed(31)
    sd = perturb_state_dict(SegformerHead(**bench.VITL_HEAD).state_dict(), seed=6)
    assert sd_digest(sd) == rec["digest"]
    g = torch.Generator().manual_seed(rec["seed"])
    feats = [torch.randn(2, 1024, 64 >> i, 64 >> i, generator=g) for i in range(4)]
    with torch.no_grad():
        out = om.segformer_head(sd, feats, prefix="")
    assert rel_l2(out, rec["out"]) < 1e-5


def test_oracle_blocks_68x120_match_reference():
    """BASELINE config 5b: Block (window: 68x120 -> 70x126 padded; global: 8160 tokens, tables interpolated 127 -> 135 / 239
    rows) and InteractionBlock(extra_extractor=True) on the 1088 x 1920 geometry."""
    from oracle import model as om
    from oracle.perturb import perturb_state_dict
    import mmsam_b200  # noqa
    from mmsam_b200 import nn_modules as M
    from mmsam_b200.ops.modules import MSDeformAttn
    rec = torch.load(os.path.join(GOLD, "blocks_68x120.pt"))
    H, W, dim, nh = rec["H"], rec["W"], rec["dim"], rec["nh"]
    x = torch.randn(1, H * W, dim, generator=torch.Generator().manual_seed(rec["x_seed"]))
    for name, ws in (("window", 14), ("global", 0)):
        torch.manual_seed(42)
        sd = perturb_state_dict(M.Block(dim, nh, 4.0, True, True, ws, (64, 64)).state_dict(), seed=7)
        r = rec[name]
        assert sd_digest(sd) == r["digest"]
        with torch.no_grad():
            out = om.vit_block(x, om.SD(sd), H, W, ws, nh)
        assert rel_l2(out.reshape(-1)[r["idx"]], r["vals"]) < 1e-5 and abs(out.norm().item() - r["norm"]) / r["norm"] < 1e-5
    r = rec["interaction"]
    torch.manual_seed(44)
    sd = perturb_state_dict(M.InteractionBlock(dim, nh, 4, True, 0.25, 0.5, 0.5, True, MSDeformAttn).state_dict(), seed=8)
    assert sd_digest(sd) == r["digest"]
    g = torch.Generator().manual_seed(r["seed"])
    xq = torch.randn(1, H * W, dim, generator=g)
    c = torch.randn(1, r["S3"], dim, generator=g)
    s = om.SD(sd)
    (ref1, sh1), (ref2, sh2) = om.deform_inputs(r["Hi"], r["Wi"], torch.float32)
    with torch.no_grad():
        xo = om.injector(xq, c, ref1, sh1, s.sub("injector"), nh, 4)
        co = om.extractor(c, xo, ref2, sh2, s.sub("extractor"), nh, 4, H, W)
        for j in range(2):
            co = om.extractor(co, xo, ref2, sh2, s.sub(f"extra_extractors.{j}"), nh, 4, H, W)
    assert rel_l2(xo.reshape(-1)[r["idx_x"]], r["vals_x"]) < 1e-5
    assert rel_l2(co.reshape(-1)[r["idx_c"]], r["vals_c"]) < 1e-5


def _full_size_oracle_check(rec, bcfg, hcfg, size, kind, btype, zero_rows_from=None):
    from oracle import model as om
    from oracle.perturb import synthetic_batch
    _, sd = build_segmentor(bcfg, hcfg, btype=btype)
    assert sd_digest(sd) == rec["digest"]
    x = synthetic_batch(1, size, kind=kind)
    if zero_rows_from is not None:
        x[:, :, zero_rows_from:] = 0
    torch.set_num_threads(max(os.cpu_count() or 1, 1))
    with torch.no_grad():
        feats = om.backbone_forward(sd, bcfg, x, prefix="backbone.")
        hl = om.segformer_head(sd, feats)
        lg = torch.nn.functional.interpolate(hl, size=x.shape[2:], mode="bilinear", align_corners=False)
        tc = rec["test_cfg"]
        if tc["mode"] == "whole_dim" or (tc["mode"] == "whole_dim_cut" and rec["rescale"]):
            lg = torch.nn.functional.interpolate(lg, size=tuple(tc["dim"]), mode="bilinear", align_corners=False)
        if tc["mode"] == "whole_dim_cut":
            lg = lg[:, :, :tc["cut_dim"][1], :tc["cut_dim"][0]]
        labels = lg.softmax(1).argmax(1)
    for f, idx, vals, nrm in zip(feats, rec["idx"], rec["vals"], rec["norms"]):
        assert rel_l2(f.reshape(-1)[idx], vals) < 1e-4
        assert abs(f.norm().item() - nrm) / nrm < 1e-4
    assert rel_l2(hl.reshape(-1)[rec["head_idx"]], rec["head_vals"]) < 1e-4
    agree = (labels == rec["labels"].long()).float().mean().item()
    print(f"oracle vs reference labels: {agree * 100:.4f}% of {labels.numel()} pixels")
    assert labels.shape == rec["labels"].shape and agree >= 0.9995


@pytest.mark.timeout(900)
def test_oracle_fmb800_config4_matches_reference():
    """BASELINE config 4 at full size: ...NEWwithcp ViT-L, 800 x 800 (tokens 50 x 50 -> windows padded to 56, rel-pos 127 ->
    99 rows), 14 classes, whole_dim_cut with rescale=False -> [600, 800] labels."""
    from common import FMB, FMB_HEAD
    rec = torch.load(os.path.join(GOLD, "fmb800_samples.pt"))
    _full_size_oracle_check(rec, FMB, FMB_HEAD, 800, "thermal", "SAMAdapterbimodalMixModNewInTwinConvNEWwithcp", 600)


@pytest.mark.timeout(900)
def test_oracle_vitl1024_config2_matches_reference():
    """BASELINE config 2 at full size (one image): ViT-L 1024 x 1024, 25 classes, whole_dim."""
    import bench
    rec = torch.load(os.path.join(GOLD, "vitl1024_samples.pt"))
    _full_size_oracle_check(rec, bench.VITL, bench.VITL_HEAD, 1024, "lidar", "SAMAdapterbimodalMixModNewInTwinConvNEW")
