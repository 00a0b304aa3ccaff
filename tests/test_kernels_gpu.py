"""GPU parity tests of the individual sm_100a kernels, called through the C ABI."""
import math

import pytest
import torch

from common import rel_l2  # noqa: E402

pytestmark = pytest.mark.gpu


def _k():
    import mmsam_b200  # noqa: F401
    from mmsam_b200 import kernels
    return kernels


def _msda_inputs(N, M, D, Lq, shapes, P, seed, spread=1.2, scale=1.0, dtype=torch.float32):
    g = torch.Generator().manual_seed(seed)
    shapes_t = torch.as_tensor(shapes, dtype=torch.long)
    S = int(shapes_t.prod(1).sum())
    L = len(shapes)
    lsi = torch.cat((shapes_t.new_zeros((1,)), shapes_t.prod(1).cumsum(0)[:-1]))
    value = (torch.rand(N, S, M, D, generator=g) * scale).to(dtype)
    loc = torch.rand(N, Lq, M, L, P, 2, generator=g) * spread - (spread - 1) / 2
    aw = torch.rand(N, Lq, M, L, P, generator=g) + 1e-5
    aw = aw / aw.sum(-1, keepdim=True).sum(-2, keepdim=True)
    return value, shapes_t, lsi, loc, aw


@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
def test_msda_reference_known_answer_shapes(dtype):
    """The reference's own check (segmentation/ops/test.py:16-75): N1 M2 D2 Lq2 L2 P2, seed 3."""
    from oracle.msda import ms_deform_attn_core
    k = _k()
    N, M, D, Lq, L, P = 1, 2, 2, 2, 2, 2
    shapes = torch.as_tensor([(6, 4), (3, 2)], dtype=torch.long)
    lsi = torch.cat((shapes.new_zeros((1,)), shapes.prod(1).cumsum(0)[:-1]))
    S = int(shapes.prod(1).sum())
    torch.manual_seed(3)
    value = torch.rand(N, S, M, D) * 0.01
    loc = torch.rand(N, Lq, M, L, P, 2)
    aw = torch.rand(N, Lq, M, L, P) + 1e-5
    aw /= aw.sum(-1, keepdim=True).sum(-2, keepdim=True)
    ref = ms_deform_attn_core(value.to(dtype), shapes, loc.to(dtype), aw.to(dtype))
    out = k.msda_forward(value.to(dtype).cuda(), shapes.cuda(), lsi.cuda(), loc.to(dtype).cuda(), aw.to(dtype).cuda()).cpu()
    if dtype == torch.float64:
        assert torch.allclose(out, ref)  # reference criterion for double (ops/test.py:44)
    else:
        assert torch.allclose(out, ref, rtol=1e-2, atol=1e-3)  # ops/test.py:68
        assert (out - ref).abs().max() < 1e-6


@pytest.mark.parametrize("vdt,adt", [(torch.bfloat16, torch.float32), (torch.float32, torch.float32),
                                     (torch.float16, torch.float16), (torch.bfloat16, torch.bfloat16)])
@pytest.mark.parametrize("cfg", [
    dict(N=2, M=16, D=32, Lq=256, shapes=[(32, 32), (16, 16), (8, 8)], P=4),    # injector-like (3 levels)
    dict(N=2, M=16, D=32, Lq=1344, shapes=[(16, 16)], P=4),                      # extractor-like (1 level)
    dict(N=1, M=12, D=32, Lq=77, shapes=[(19, 25), (10, 13)], P=4),              # non-square, ragged Lq
    dict(N=1, M=3, D=24, Lq=5, shapes=[(7, 5)], P=2),                            # odd head count
    dict(N=1, M=2, D=7, Lq=9, shapes=[(4, 4), (2, 2)], P=3),                     # generic path
])
def test_msda_vs_oracle(cfg, vdt, adt):
    from oracle.msda import ms_deform_attn_core
    k = _k()
    value, shapes, lsi, loc, aw = _msda_inputs(seed=1, **cfg)
    v = value.to(vdt)
    lc, w = loc.to(adt), aw.to(adt)
    ref = ms_deform_attn_core(v.double(), shapes, lc.double(), w.double())
    out = k.msda_forward(v.cuda(), shapes.cuda(), lsi.cuda(), lc.cuda(), w.cuda()).cpu().double()
    err = (out - ref).abs().max().item()
    # north_star: within 1e-3 of ms_deform_attn_core_pytorch (same rounded inputs, fp32 accumulate);
    # bf16/fp16 outputs add one rounding of an O(1) value
    tol = 1e-3 if vdt == torch.float32 else (8e-3 if vdt == torch.bfloat16 else 2e-3)
    assert err < tol, err


def test_msda_unit_scale_and_small_scale():
    """ops/test.py scales value by 0.01; also check at unit scale with rtol 1e-2 / atol 1e-3."""
    from oracle.msda import ms_deform_attn_core
    k = _k()
    for scale in (0.01, 1.0):
        value, shapes, lsi, loc, aw = _msda_inputs(N=1, M=16, D=32, Lq=512, shapes=[(24, 24), (12, 12), (6, 6)], P=4,
                                                   seed=5, scale=scale)
        v = value.to(torch.bfloat16)
        ref = ms_deform_attn_core(v.float(), shapes, loc, aw)
        out = k.msda_forward(v.cuda(), shapes.cuda(), lsi.cuda(), loc.cuda(), aw.cuda()).cpu().float()
        assert torch.allclose(out, ref, rtol=1e-2, atol=1e-3), (out - ref).abs().max()
        if scale == 0.01:
            assert (out - ref).abs().max() < 1e-3


@pytest.mark.parametrize("N,M,Lq,shapes,spread", [
    (2, 16, 256, [(32, 32), (16, 16), (8, 8)], 3.0),      # injector-like: 3 levels
    (2, 16, 1344, [(16, 16)], 3.0),                        # extractor-like: 1 level
    (1, 12, 77, [(19, 25), (10, 13)], 6.0),                # ViT-B head count (not a multiple of 8), non-square, ragged
    (1, 16, 50, [(9, 7)], 40.0),                           # most samples outside the map
])
def test_msda_fused_vs_oracle(N, M, Lq, shapes, spread):
    """Fused entry point (softmax over L*P logits + loc = ref + off / (W, H) in the kernel,
    ops/modules/ms_deform_attn.py:108-127) against the oracle fed the materialised locations / weights."""
    from oracle.msda import ms_deform_attn_core
    k = _k()
    D, P = 32, 4
    L = len(shapes)
    g = torch.Generator().manual_seed(N * 1000 + M * 10 + Lq)
    shapes_t = torch.as_tensor(shapes, dtype=torch.long)
    S = int(shapes_t.prod(1).sum())
    lsi = torch.cat((shapes_t.new_zeros((1,)), shapes_t.prod(1).cumsum(0)[:-1]))
    value = torch.randn(N, S, M * D, generator=g).to(torch.bfloat16)
    ref = torch.rand(Lq, 2, generator=g)
    pad = 6                                                # qproj rows carry extra columns (ldq > M*L*P*3, even)
    qproj = torch.randn(N * Lq, M * L * P * 3 + pad, generator=g)
    qproj[:, :M * L * P * 2] *= spread
    off = qproj[:, :M * L * P * 2].view(N, Lq, M, L, P, 2)
    logits = qproj[:, M * L * P * 2:M * L * P * 3].view(N, Lq, M, L * P)
    norm = torch.stack([shapes_t[:, 1], shapes_t[:, 0]], -1).float()          # (W, H)
    loc = ref[None, :, None, None, None, :] + off / norm[None, None, None, :, None, :]
    aw = torch.softmax(logits, -1).view(N, Lq, M, L, P)
    exp = ms_deform_attn_core(value.float().view(N, S, M, D), shapes_t, loc, aw)
    out = k.msda_fused(value.cuda(), shapes_t.cuda(), lsi.cuda(), qproj.cuda(), ref.cuda(), M, L, P).cpu().float()
    assert (out - exp).abs().max() < 1e-3 + 0.008 * exp.abs().max()          # one bf16 rounding of the output


def _grid_refs(qgrids):
    refs = []
    for (h, w) in qgrids:
        ys = torch.linspace(0.5, h - 0.5, h) / h
        xs = torch.linspace(0.5, w - 0.5, w) / w
        yy, xx = torch.meshgrid(ys, xs, indexing="ij")
        refs.append(torch.stack((xx.reshape(-1), yy.reshape(-1)), -1))
    return torch.cat(refs).float().contiguous()


@pytest.mark.parametrize("N,M,levels,qgrids,anchor,tile,noise,margin", [
    (2, 16, [(32, 32), (16, 16), (8, 8)], [(16, 16)], (16, 16), (8, 16), 0.3, 2),        # injector-like
    (2, 16, [(16, 16)], [(32, 32), (16, 16), (8, 8)], (16, 16), (8, 16), 0.3, 2),        # extractor-like
    (1, 12, [(40, 56), (20, 28), (10, 14)], [(20, 28)], (20, 28), (8, 16), 0.5, 2),      # 320x448 geometry, ViT-B heads, ragged tiles
    (1, 12, [(20, 28)], [(40, 56), (20, 28), (10, 14)], (20, 28), (8, 16), 0.5, 2),
    (1, 16, [(25, 25)], [(50, 50), (25, 25), (12, 12)], (25, 25), (8, 16), 0.3, 1),      # FMB-like: 12 = 100 // 8, non-integer grid ratios
    (1, 16, [(16, 16), (8, 8)], [(16, 16)], (16, 16), (4, 4), 6.0, 1),                   # offsets far from the prior: global fall-back path
    (1, 16, [(9, 7)], [(9, 7)], (9, 7), (8, 16), 40.0, 0),                               # most samples outside the map
])
def test_msda_staged_vs_oracle(N, M, levels, qgrids, anchor, tile, noise, margin):
    """Shared-memory staged kernel (TMA boxes placed by the sampling_offsets-bias prior, global fall-back for samples
    that leave their box) against the oracle fed the materialised locations / weights."""
    from oracle.msda import ms_deform_attn_core
    k = _k()
    D, P = 32, 4
    L = len(levels)
    g = torch.Generator().manual_seed(N * 1000 + M * 10 + L)
    shapes_t = torch.as_tensor(levels, dtype=torch.long)
    S = int(shapes_t.prod(1).sum())
    lsi = torch.cat((shapes_t.new_zeros((1,)), shapes_t.prod(1).cumsum(0)[:-1]))
    ref = _grid_refs(qgrids)
    Lq = ref.shape[0]
    value = torch.randn(N, S, M * D, generator=g).to(torch.bfloat16)
    # MSDeformAttn._reset_parameters: head direction x (p + 1) pixels (ops/modules/ms_deform_attn.py:64-74)
    th = torch.arange(M).float() * (2 * math.pi / M)
    gi = torch.stack([th.cos(), th.sin()], -1)
    gi = gi / gi.abs().max(-1, keepdim=True)[0]
    bias = (gi.view(M, 1, 1, 2) * torch.arange(1, P + 1).view(1, 1, P, 1).float()).expand(M, L, P, 2).contiguous()
    pad = 8                                                # the staged kernel copies 16-byte pieces: ldq % 4 == 0
    qproj = torch.randn(N * Lq, M * L * P * 3 + pad, generator=g)
    qproj[:, :M * L * P * 2] = qproj[:, :M * L * P * 2] * noise + bias.reshape(-1)
    off = qproj[:, :M * L * P * 2].view(N, Lq, M, L, P, 2)
    logits = qproj[:, M * L * P * 2:M * L * P * 3].view(N, Lq, M, L * P)
    norm = torch.stack([shapes_t[:, 1], shapes_t[:, 0]], -1).float()
    loc = ref[None, :, None, None, None, :] + off / norm[None, None, None, :, None, :]
    aw = torch.softmax(logits, -1).view(N, Lq, M, L, P)
    exp = ms_deform_attn_core(value.float().view(N, S, M, D), shapes_t, loc, aw)
    geom = k.MsdaGeometry(levels, qgrids, anchor, tile, bias, M, L, P, margin=margin)
    out = torch.full((N, Lq, M * D), float("nan"), dtype=torch.bfloat16, device="cuda")   # every query must be written
    k.msda_fused(value.cuda(), shapes_t.cuda(), lsi.cuda(), qproj.cuda(), ref.cuda(), M, L, P, out=out, geom=geom)
    assert not geom.unsupported
    out = out.cpu().float()
    assert torch.isfinite(out).all()
    assert (out - exp).abs().max() < 1e-3 + 0.008 * exp.abs().max()
    # and the staged kernel agrees with the L1-gather kernel to fp32 summation order
    out2 = k.msda_fused(value.cuda(), shapes_t.cuda(), lsi.cuda(), qproj.cuda(), ref.cuda(), M, L, P).cpu().float()
    assert (out - out2).abs().max() <= 0.008 * exp.abs().max() + 1e-6


@pytest.mark.parametrize("N,M,D,Lq,levels,P,dtype,spread", [
    (1, 2, 2, 2, [(6, 4), (3, 2)], 2, torch.float64, 1.0),      # the reference's own check shape (ops/test.py:16-75)
    (2, 8, 32, 37, [(16, 12), (8, 6), (4, 3)], 4, torch.float64, 1.0),
    (1, 3, 48, 20, [(9, 7)], 4, torch.float64, 1.6),            # D > 32 (strided lanes), samples outside the map
    (2, 4, 16, 33, [(10, 14), (5, 7)], 4, torch.float32, 1.2),
])
def test_msda_backward_vs_autograd_of_the_oracle(N, M, D, Lq, levels, P, dtype, spread):
    """mmsam_msda_backward (MSDA.ms_deform_attn_backward, ops/src/cuda/ms_deform_attn_cuda.cu:83-153) against autograd
    through the oracle's forward (the arithmetic of ms_deform_attn_core_pytorch / the reference kernel)."""
    from oracle.msda import ms_deform_attn_core
    k = _k()
    g = torch.Generator().manual_seed(N * 100 + M * 10 + D)
    L = len(levels)
    shapes_t = torch.as_tensor(levels, dtype=torch.long)
    S = int(shapes_t.prod(1).sum())
    lsi = torch.cat((shapes_t.new_zeros((1,)), shapes_t.prod(1).cumsum(0)[:-1]))
    value = torch.randn(N, S, M, D, generator=g, dtype=dtype).requires_grad_()
    loc = ((torch.rand(N, Lq, M, L, P, 2, generator=g, dtype=dtype) - 0.5) * spread + 0.5).requires_grad_()
    aw = torch.rand(N, Lq, M, L, P, generator=g, dtype=dtype) + 1e-5
    aw = (aw / aw.sum((-1, -2), keepdim=True)).requires_grad_()
    go = torch.randn(N, Lq, M * D, generator=g, dtype=dtype)
    ms_deform_attn_core(value, shapes_t, loc, aw).backward(go)
    gv, gl, ga = k.msda_backward(value.detach().cuda(), shapes_t.cuda(), lsi.cuda(), loc.detach().cuda(), aw.detach().cuda(), go.cuda())
    tol = 1e-9 if dtype == torch.float64 else 2e-4
    for name, got, want in (("grad_value", gv, value.grad), ("grad_sampling_loc", gl, loc.grad), ("grad_attn_weight", ga, aw.grad)):
        err = (got.cpu() - want).abs().max().item()
        assert err <= tol * max(1.0, want.abs().max().item()), (name, err)


def test_msda_function_numerical_gradcheck():
    """The reference's check_gradient_numerical (ops/test.py:55-75): torch.autograd.gradcheck of MSDeformAttnFunction in
    double, every differentiable input."""
    import mmsam_b200  # noqa
    from mmsam_b200.ops.functions import MSDeformAttnFunction
    g = torch.Generator().manual_seed(3)
    N, M, D, Lq, L, P = 1, 2, 4, 2, 2, 2
    shapes = torch.as_tensor([(6, 4), (3, 2)], dtype=torch.long).cuda()
    lsi = torch.cat((shapes.new_zeros((1,)), shapes.prod(1).cumsum(0)[:-1]))
    S = int(shapes.prod(1).sum())
    value = (torch.rand(N, S, M, D, generator=g, dtype=torch.float64) * 0.01).cuda().requires_grad_()
    loc = torch.rand(N, Lq, M, L, P, 2, generator=g, dtype=torch.float64).cuda().requires_grad_()
    aw = torch.rand(N, Lq, M, L, P, generator=g, dtype=torch.float64) + 1e-5
    aw = (aw / aw.sum((-1, -2), keepdim=True)).cuda().requires_grad_()
    assert torch.autograd.gradcheck(MSDeformAttnFunction.apply, (value, shapes, lsi, loc, aw, 2), nondet_tol=1e-12)


def test_msda_empty():
    k = _k()
    shapes = torch.as_tensor([(4, 4)], dtype=torch.long).cuda()
    lsi = torch.zeros(1, dtype=torch.long).cuda()
    v = torch.zeros(1, 16, 2, 8, device="cuda", dtype=torch.bfloat16)
    out = k.msda_forward(v, shapes, lsi, torch.zeros(1, 0, 2, 1, 4, 2, device="cuda"), torch.zeros(1, 0, 2, 1, 4, device="cuda"))
    assert out.shape == (1, 0, 16)


@pytest.mark.parametrize("C", [96, 192, 384, 768, 1024, 1536, 64])
def test_layernorm(C):
    k = _k()
    g = torch.Generator().manual_seed(C)
    x = (torch.randn(1000, C, generator=g) * 2 + 0.5).to(torch.bfloat16)
    w = 1 + 0.1 * torch.randn(C, generator=g)
    b = 0.1 * torch.randn(C, generator=g)
    ref = torch.nn.functional.layer_norm(x.float(), (C,), w, b, 1e-6)
    out = k.layernorm(x.cuda(), w.cuda(), b.cuda(), 1e-6).cpu().float()
    assert (out - ref).abs().max() < 0.04  # one bf16 rounding of |y| <~ 5
    assert ((out - ref).abs() <= 0.004 * ref.abs() + 1e-3).all()


def test_layernorm_row_map():
    k = _k()
    C = 128
    x = torch.randn(10, C).to(torch.bfloat16)
    rm = torch.tensor([3, -1, 0, 5, 11, -1, 7, 8, 1, 2], dtype=torch.int32)
    w, b = torch.ones(C), torch.zeros(C)
    out = k.layernorm(x.cuda(), w.cuda(), b.cuda(), 1e-6, row_map=rm.cuda(), out_rows=12).cpu().float()
    ref = torch.nn.functional.layer_norm(x.float(), (C,), w, b, 1e-6)
    exp = torch.zeros(12, C)
    for i, d in enumerate(rm.tolist()):
        if d >= 0:
            exp[d] = ref[i]
    assert (out - exp).abs().max() < 0.04


def _gemm_ref(a, w, bias=None, act=None, scale=None, residual=None):
    y = a.double() @ w.double().t()
    if bias is not None:
        y = y + bias.double()
    if act == "gelu":
        y = torch.nn.functional.gelu(y)
    elif act == "relu":
        y = torch.relu(y)
    elif act == "relu6":
        y = torch.clamp(y, 0, 6)
    if scale is not None:
        y = y * scale.double()
    if residual is not None:
        y = y + residual.double()
    return y


@pytest.mark.parametrize("M,N,K,bn", [
    (128, 256, 64, 256), (128, 256, 256, 256), (256, 512, 1024, 256), (300, 200, 96, 0), (128, 128, 128, 128),
    (1000, 64, 768, 64), (4096, 3072, 1024, 0), (4900, 1024, 1024, 0), (77, 25, 512, 0), (513, 384, 48, 0),
    (20000, 4096, 1024, 256), (640, 1024, 4096, 128),
])
def test_gemm_plain(M, N, K, bn):
    k = _k()
    g = torch.Generator().manual_seed(M + N + K)
    a = torch.randn(M, K, generator=g).to(torch.bfloat16)
    w = (torch.randn(N, K, generator=g) / math.sqrt(K)).to(torch.bfloat16)
    bias = torch.randn(N, generator=g)
    out = k.gemm(a.cuda(), w.cuda(), bias=bias.cuda(), block_n=bn, out_dtype=torch.float32).cpu().double()
    ref = _gemm_ref(a, w, bias)
    err = (out - ref).abs().max().item()
    assert err < 2e-3, err
    if N % 8 == 0:
        outb = k.gemm(a.cuda(), w.cuda(), bias=bias.cuda(), block_n=bn).cpu().double()
        assert ((outb - ref).abs() <= 0.008 * ref.abs() + 2e-3).all()


@pytest.mark.parametrize("act", ["gelu", "relu", "relu6", None])
def test_gemm_epilogues(act):
    k = _k()
    M, N, K = 777, 512, 256
    g = torch.Generator().manual_seed(11)
    a = torch.randn(M, K, generator=g).to(torch.bfloat16)
    w = (torch.randn(N, K, generator=g) / math.sqrt(K) * 3).to(torch.bfloat16)
    bias = torch.randn(N, generator=g)
    scale = torch.randn(N, generator=g)
    res = torch.randn(M, N, generator=g).to(torch.bfloat16)
    out = k.gemm(a.cuda(), w.cuda(), bias=bias.cuda(), act=act, scale=scale.cuda(), residual=res.cuda(),
                 out_dtype=torch.float32).cpu().double()
    ref = _gemm_ref(a, w, bias, act, scale, res)
    assert (out - ref).abs().max() < 3e-3


def test_gemm_strided_operands_and_row_map():
    k = _k()
    M, N, K = 392, 256, 128
    g = torch.Generator().manual_seed(5)
    big = torch.randn(M, 3 * K, generator=g).to(torch.bfloat16).cuda()
    a = big[:, K:2 * K]                      # lda = 3K
    w = (torch.randn(N, K, generator=g) / 8).to(torch.bfloat16).cuda()
    rm = torch.randperm(M + 50, generator=g)[:M].to(torch.int32)
    rm[::7] = -1
    res = torch.randn(M + 50, N, generator=g).to(torch.bfloat16)
    out = torch.full((M + 50, N), 7.0, dtype=torch.bfloat16, device="cuda")
    k.gemm(a, w, residual=res.cuda(), out=out, row_map=rm.cuda())
    ref = _gemm_ref(a.cpu(), w.cpu())
    exp = torch.full((M + 50, N), 7.0, dtype=torch.float64)
    for i, d in enumerate(rm.tolist()):
        if d >= 0:
            exp[d] = ref[i] + res[d].double()
    assert (out.cpu().double() - exp).abs().max() < 0.05


def test_gemm_pixel_shuffle():
    """ConvTranspose2d(C, C, 2, 2) as one GEMM + pixel-shuffle store (backbone `up`, ..._new.py:324)."""
    k = _k()
    B, H, W, Cin, Cout = 2, 6, 5, 64, 32
    g = torch.Generator().manual_seed(2)
    x = torch.randn(B, H, W, Cin, generator=g).to(torch.bfloat16)
    wt = (torch.randn(Cin, Cout, 2, 2, generator=g) / 8).to(torch.bfloat16)
    bias = torch.randn(Cout, generator=g)
    ref = torch.nn.functional.conv_transpose2d(x.float().permute(0, 3, 1, 2), wt.float(), bias, stride=2)
    ref = ref.permute(0, 2, 3, 1).reshape(-1, Cout)
    wg = wt.permute(2, 3, 1, 0).reshape(4 * Cout, Cin).contiguous()
    out = k.gemm(x.reshape(-1, Cin).cuda(), wg.cuda(), bias=bias.repeat(4).cuda(), pixel_shuffle=(H, W),
                 out_dtype=torch.float32).cpu()
    assert (out - ref).abs().max() < 2e-3


@pytest.mark.parametrize("M,N,K,bn", [
    (1000, 384, 1536, 256),    # last n-tile half empty: some epilogue warps have no panel in odd tiles
    (700, 1024, 256, 0), (4173, 384, 64, 128), (300, 64, 128, 64), (520, 1024, 4096, 256), (33000, 384, 96, 0),
    (20000, 1024, 512, 0),
])
def test_gemm_bf16_residual_paths(M, N, K, bn):
    """bf16 output + bias + per-channel scale + residual through both residual pipelines (cp.async ring for
    K <= 2048, register prefetch above), ragged M / N tiles, every tile width."""
    k = _k()
    g = torch.Generator().manual_seed(M * 3 + N + K)
    a = torch.randn(M, K, generator=g).to(torch.bfloat16)
    w = (torch.randn(N, K, generator=g) / math.sqrt(K)).to(torch.bfloat16)
    bias = torch.randn(N, generator=g)
    scale = torch.rand(N, generator=g) + 0.5
    res = torch.randn(M, N, generator=g).to(torch.bfloat16)
    out = k.gemm(a.cuda(), w.cuda(), bias=bias.cuda(), scale=scale.cuda(), residual=res.cuda(), block_n=bn).cpu().double()
    ref = _gemm_ref(a, w, bias, None, scale, res)
    assert ((out - ref).abs() <= 0.008 * ref.abs() + 4e-3).all(), (out - ref).abs().max().item()


def test_gemm_pixel_shuffle_bf16_residual():
    """`up(c2) + c1` (..._new.py:324-325) as one GEMM: pixel-shuffle store with the residual read at the destination."""
    k = _k()
    B, H, W, Cin, Cout = 2, 16, 12, 128, 64
    g = torch.Generator().manual_seed(4)
    x = torch.randn(B, H, W, Cin, generator=g).to(torch.bfloat16)
    wt = (torch.randn(Cin, Cout, 2, 2, generator=g) / 8).to(torch.bfloat16)
    bias = torch.randn(Cout, generator=g)
    res = torch.randn(B * 4 * H * W, Cout, generator=g).to(torch.bfloat16)
    ref = torch.nn.functional.conv_transpose2d(x.float().permute(0, 3, 1, 2), wt.float(), bias, stride=2)
    ref = ref.permute(0, 2, 3, 1).reshape(-1, Cout).double() + res.double()
    wg = wt.permute(2, 3, 1, 0).reshape(4 * Cout, Cin).contiguous()
    out = k.gemm(x.reshape(-1, Cin).cuda(), wg.cuda(), bias=bias.repeat(4).cuda(), residual=res.cuda(),
                 pixel_shuffle=(H, W)).cpu().double()
    assert ((out - ref).abs() <= 0.008 * ref.abs() + 4e-3).all()


def test_gemm_row_map_big():
    """Row-mapped store (c2|c3|c4 packing, window un-partition) over many tiles, with and without residual."""
    k = _k()
    M, N, K = 5000, 512, 192
    g = torch.Generator().manual_seed(6)
    a = torch.randn(M, K, generator=g).to(torch.bfloat16)
    w = (torch.randn(N, K, generator=g) / math.sqrt(K)).to(torch.bfloat16)
    rm = torch.randperm(M + 300, generator=g)[:M].to(torch.int32)
    rm[::11] = -1
    res = torch.randn(M + 300, N, generator=g).to(torch.bfloat16)
    ref = _gemm_ref(a, w)
    for use_res in (False, True):
        out = torch.full((M + 300, N), 3.0, dtype=torch.bfloat16, device="cuda")
        k.gemm(a.cuda(), w.cuda(), residual=res.cuda() if use_res else None, out=out, row_map=rm.cuda())
        exp = torch.full((M + 300, N), 3.0, dtype=torch.float64)
        keep = rm >= 0
        exp[rm[keep].long()] = ref[keep] + (res.double()[rm[keep].long()] if use_res else 0)
        assert ((out.cpu().double() - exp).abs() <= 0.008 * exp.abs() + 4e-3).all()


@pytest.mark.parametrize("Bp,nh,Kh,Kw,bias", [
    (2, 2, 14, 14, True),     # SAM window (196 keys, ragged 2nd key block)
    (1, 2, 16, 16, True),     # 256 tokens, two full key blocks
    (1, 1, 8, 16, False),     # single block, no bias
    (1, 2, 32, 32, True),     # ViT-B/512 global (1024 tokens)
    (2, 3, 19, 25, True),     # non-square grid (475 tokens, ragged)
    (3, 4, 14, 14, False),    # SAM window without bias (single-pass window kernel)
    (150, 16, 14, 14, True),  # more (window, head) items than CTAs x 2 stages: the window kernel's ring wraps several times
    (50, 16, 14, 14, True),   # many windows: exercises the persistent tile loop (2 images x 25 windows)
    (1, 16, 64, 64, True),    # ViT-L/1024 global: 4096 tokens, 127-row tables, two G chunks
    (1, 2, 68, 120, True),    # MUSES whole frame 1088x1920 (BASELINE config 5b): 8160 tokens, half-precision bias rows
])
def test_attention_vs_oracle(Bp, nh, Kh, Kw, bias):
    from oracle.model import attention_core
    k = _k()
    T = Kh * Kw
    g = torch.Generator().manual_seed(Bp * 1000 + T)
    qkv = (torch.randn(Bp, T, 3, nh, 64, generator=g) * 1.5).to(torch.bfloat16)
    rph = (torch.randn(2 * Kh - 1, 64, generator=g) * 0.2).to(torch.bfloat16)
    rpw = (torch.randn(2 * Kw - 1, 64, generator=g) * 0.2).to(torch.bfloat16)
    th = tw = None
    if bias:
        th, tw = k.relpos_table(rph.cuda(), Kh), k.relpos_table(rpw.cuda(), Kw)
    out = k.attention(qkv.view(Bp, T, -1).cuda(), nh, (Kh, Kw), th, tw).cpu().float()
    q, kk, v = qkv.float().permute(2, 0, 3, 1, 4).reshape(3, Bp * nh, T, 64).unbind(0)
    nchk = min(Bp * nh, 48)
    dt = torch.float64 if T <= 4096 else torch.float32      # the 8160-token score matrix is 0.5 GB in double
    ref = attention_core(q[:nchk].to(dt), kk[:nchk].to(dt), v[:nchk].to(dt), Kh, Kw,
                         rph.to(dt) if bias else None, rpw.to(dt) if bias else None).float()
    got = out.view(Bp, T, nh, 64).permute(0, 2, 1, 3).reshape(Bp * nh, T, 64)[:nchk]
    err = (got - ref).abs().max().item()
    rel = ((got - ref).norm() / ref.norm()).item()
    assert rel < 1e-2 and err < 0.05, (rel, err)


@pytest.mark.parametrize("Bp,nh,Kh,ramp", [
    (1, 2, 2, 0.0),      # 128 tokens: one key block, tile B of the only item has no rows
    (2, 3, 6, 0.0),      # 384 tokens: the second item of a (batch, head) has only tile A
    (1, 2, 34, 0.0),     # 2176 tokens = 17 tiles, 65-row key-row table (zero-filled to 128 by the copy engine)
    (5, 16, 8, 0.0),     # more items than CTAs x ...: 160 items over 148 CTAs, the persistent loop wraps
    (1, 2, 16, 24.0),    # logits ramp by ~24 log2 units per key block: every block takes the power-of-two rescale path
])
def test_attention_global_two_tile_kernel(Bp, nh, Kh, ramp):
    """64-wide token grids run the two-tile ping-pong kernel (attention_glb.cu): single-pass softmax with a stale reference
    maximum, P in tensor memory, key-column bias in registers. Against the fp64 oracle (base/image_encoder.py:483-501,
    587-623), including key rows whose logits keep growing (the exact power-of-two rescale of P, sum and accumulator)."""
    from oracle.model import attention_core
    k = _k()
    Kw, T = 64, Kh * 64
    g = torch.Generator().manual_seed(Bp * 1000 + T)
    qkv = (torch.randn(Bp, T, 3, nh, 64, generator=g) * 1.5)
    if ramp:
        # k = c(key row) * q_dir: the score of every query with the later key rows grows without bound along the row
        qdir = torch.randn(64, generator=g)
        qdir = qdir / qdir.norm()
        qkv[:, :, 0] = qkv[:, :, 0] * 0.2 + 3.0 * qdir
        rows = torch.arange(T) // 128
        qkv[:, :, 1] = qkv[:, :, 1] * 0.2 + (ramp * 8 * 0.6931 / 3.0) * rows[None, :, None, None].float() * qdir
    qkv = qkv.to(torch.bfloat16)
    rph = (torch.randn(2 * Kh - 1, 64, generator=g) * 0.2).to(torch.bfloat16)
    rpw = (torch.randn(2 * Kw - 1, 64, generator=g) * 0.2).to(torch.bfloat16)
    th, tw = k.relpos_table(rph.cuda(), Kh), k.relpos_table(rpw.cuda(), Kw)
    out = k.attention(qkv.view(Bp, T, -1).cuda(), nh, (Kh, Kw), th, tw).cpu().float()
    q, kk, v = qkv.float().permute(2, 0, 3, 1, 4).reshape(3, Bp * nh, T, 64).unbind(0)
    nchk = min(Bp * nh, 48)
    ref = attention_core(q[:nchk].double(), kk[:nchk].double(), v[:nchk].double(), Kh, Kw, rph.double(), rpw.double()).float()
    got = out.view(Bp, T, nh, 64).permute(0, 2, 1, 3).reshape(Bp * nh, T, 64)[:nchk]
    assert torch.isfinite(got).all()
    err = (got - ref).abs().max().item()
    rel = ((got - ref).norm() / ref.norm()).item()
    assert rel < 1e-2 and err < 0.05, (rel, err)


@pytest.mark.parametrize("B,nh,H,W,bias", [(2, 16, 64, 64, True), (1, 2, 19, 25, True), (3, 3, 28, 28, True), (1, 2, 14, 70, False),
                                           (8, 16, 64, 64, True)])
def test_attention_window_unpartitioned_store(B, nh, H, W, bias):
    """SAM window attention with window_unpartition fused into the store (mmsam_attention_window_bf16: 126 | 70-row tiles,
    one 4-D TMA store per tile, clipped at the image edge) against the row-mapped path of mmsam_attention_bf16, which the
    oracle tests pin: same per-row arithmetic -> identical bits; padded window tokens must not reach the output."""
    from mmsam_b200.engine import window_maps
    k = _k()
    g = torch.Generator().manual_seed(B * 100 + H + W)
    sc = window_maps(B, H, W, 14, 8, "cuda")
    Bp = sc["win_bp"]
    qkv = (torch.randn(Bp, 196, 3 * nh * 64, generator=g) * 1.5).to(torch.bfloat16).cuda()
    th = tw = None
    if bias:
        th = k.relpos_table((torch.randn(27, 64, generator=g) * 0.2).cuda(), 14)
        tw = k.relpos_table((torch.randn(27, 64, generator=g) * 0.2).cuda(), 14)
    want = k.attention(qkv, nh, (14, 14), th, tw, out_map=sc["win_inv"], out_rows=B * H * W)
    out = torch.full((B * H * W + 64, nh * 64), 7.0, dtype=torch.bfloat16, device="cuda")      # guard rows behind the map
    k.attention_window(qkv, nh, B, H, W, th, tw, out=out[:B * H * W])
    torch.cuda.synchronize()
    assert torch.equal(out[:B * H * W], want)
    assert (out[B * H * W:] == 7.0).all()


def test_attention_relpos_interpolated_table():
    """FMB-shaped case: a 127-row table interpolated to 2*50-1 = 99 rows (get_rel_pos :566-575)."""
    from oracle.model import attention_core
    k = _k()
    Kh = Kw = 20
    T = Kh * Kw
    g = torch.Generator().manual_seed(3)
    qkv = torch.randn(1, T, 3, 2, 64, generator=g).to(torch.bfloat16)
    rph = (torch.randn(127, 64, generator=g) * 0.2)
    rpw = (torch.randn(127, 64, generator=g) * 0.2)
    th, tw = k.relpos_table(rph.cuda(), Kh), k.relpos_table(rpw.cuda(), Kw)
    out = k.attention(qkv.view(1, T, -1).cuda(), 2, (Kh, Kw), th, tw).cpu().float()
    q, kk, v = qkv.float().permute(2, 0, 3, 1, 4).reshape(3, 2, T, 64).unbind(0)
    ref = attention_core(q, kk, v, Kh, Kw, rph, rpw)
    got = out.view(1, T, 2, 64).permute(0, 2, 1, 3).reshape(2, T, 64)
    assert ((got - ref).norm() / ref.norm()).item() < 1.5e-2


@pytest.mark.parametrize("B,H,W,Cin,groups", [
    (2, 16, 16, 288, 32),   # GFE qkv2 at level 0 (cg = 9: 54-channel n-tiles, one k-block)
    (1, 16, 16, 576, 32),   # level 1 (cg = 18: 54-channel n-tiles)
    (1, 8, 16, 1152, 32),   # level 2 (cg = 36: 36-channel n-tiles)
    (1, 8, 24, 64, 32),     # cg = 2 (Mlp dwconv style), one k-block
    (1, 4, 4, 192, 96),     # map smaller than the 8x16 tile, 2 ch / group
    (1, 13, 9, 2304, 32),   # cg = 72, three k-blocks, ragged map
    (2, 32, 32, 128, 1),    # dense 3x3
])
def test_conv3x3_grouped(B, H, W, Cin, groups):
    k = _k()
    g = torch.Generator().manual_seed(Cin + groups)
    x = torch.randn(B, H, W, Cin, generator=g).to(torch.bfloat16)
    w = (torch.randn(Cin, Cin // groups, 3, 3, generator=g) / math.sqrt(9 * Cin // groups)).to(torch.bfloat16)
    wp = k.pack_conv3x3_weight(w.float().cuda(), groups)
    out = k.conv3x3(x.reshape(-1, Cin).cuda(), wp, B, H, W, Cin, Cin, groups).float().cpu().view(B, H, W, Cin)
    ref = torch.nn.functional.conv2d(x.float().permute(0, 3, 1, 2), w.float(), padding=1, groups=groups).permute(0, 2, 3, 1)
    assert ((out - ref).abs() <= 0.01 * ref.abs() + 5e-3).all(), (out - ref).abs().max()


@pytest.mark.parametrize("n,blk,HW,norms", [
    (96, 12, 4096, True), (64, 0, 1000, True), (192, 0, 300, True), (768, 96, 256, True),   # SIMT fallback when HW % 64 != 0
    (192, 0, 8192, False), (384, 48, 2048, True), (384, 0, 1024, False), (1536, 0, 1024, False), (200, 0, 640, True),
])
def test_gram(n, blk, HW, norms):
    """Gram over pixels (tcgen05 path with MN-major operands for HW % 64 == 0, SIMT kernel otherwise)."""
    k = _k()
    B, ld = 2, 3 * n
    g = torch.Generator().manual_seed(n + HW)
    x = torch.randn(B, HW, ld, generator=g).to(torch.bfloat16)
    xc = x.reshape(-1, ld).cuda()
    r = k.gram(xc, ld, 0, n, n, B, HW, blk=blk, norms=norms)
    Sp = r[0] if norms else r                              # [nchunks, B, n, n] partial sums of the pixel chunks
    q, kk = x[..., :n].double(), x[..., n:2 * n].double()
    ref = q.transpose(1, 2) @ kk
    m = torch.ones(n, n, dtype=torch.bool)
    if blk:                                                # only the block diagonal is written
        m = (torch.arange(n)[:, None] // blk) == (torch.arange(n)[None, :] // blk)
    S = torch.where(m, Sp.cpu().double(), torch.zeros((), dtype=torch.float64)).sum(0)
    assert (S - ref * m).abs().max() < 2e-3 * math.sqrt(HW)
    if norms:
        nq, nk = r[1].cpu().double().sum(0), r[2].cpu().double().sum(0)
        assert torch.allclose(nq, (q * q).sum(1), rtol=1e-4) and torch.allclose(nk, (kk * kk).sum(1), rtol=1e-4)
    # no atomics: a second launch reproduces every partial sum bit for bit
    r2 = k.gram(xc, ld, 0, n, n, B, HW, blk=blk, norms=norms)
    Sp2 = r2[0] if norms else r2
    assert torch.equal(torch.where(m.cuda(), Sp, torch.zeros((), device="cuda")), torch.where(m.cuda(), Sp2, torch.zeros((), device="cuda")))


@pytest.mark.parametrize("B,C,hs,ws,ho,wo", [
    (2, 1024, 8, 8, 32, 32),     # x4 up (f1 fusion), channel vectors divide the grid stride: scale/shift in registers
    (2, 96, 16, 12, 8, 6),       # x0.5 down, 12 channel vectors: in-loop scale/shift path
    (1, 64, 5, 7, 5, 7),         # same size (f3: plain add + BatchNorm)
    (3, 512, 4, 4, 16, 16),
])
def test_resize_add_affine(B, C, hs, ws, ho, wo):
    """(base + bilinear(src)) * scale + shift, channels-last, against F.interpolate(align_corners=False) + eval-BN
    folded to scale/shift (backbone tail ..._new.py:326-337, head resize segformer_head.py:55-61)."""
    k = _k()
    g = torch.Generator().manual_seed(B * C + hs)
    src = torch.randn(B, hs, ws, C, generator=g).to(torch.bfloat16)
    base = torch.randn(B, ho, wo, C, generator=g).to(torch.bfloat16)
    scale = torch.rand(C, generator=g) + 0.5
    shift = torch.randn(C, generator=g)
    up = torch.nn.functional.interpolate(src.float().permute(0, 3, 1, 2), size=(ho, wo), mode="bilinear",
                                         align_corners=False).permute(0, 2, 3, 1)
    ref = (up + base.float()) * scale + shift
    out = k.resize_add_affine(src.cuda(), (hs, ws), (ho, wo), B, C, base=base.cuda(), scale=scale.cuda(),
                              shift=shift.cuda()).cpu().float()
    assert ((out - ref).abs() <= 0.008 * ref.abs() + 4e-3).all()
    out2 = k.resize_add_affine(src.cuda(), (hs, ws), (ho, wo), B, C).cpu().float()       # plain resize
    assert ((out2 - up).abs() <= 0.008 * up.abs() + 4e-3).all()


@pytest.mark.parametrize("M,N,K,f32,act", [(1000, 192, 1024, True, None), (4096, 512, 1024, False, None),
                                             (777, 256, 1024, False, "gelu"), (300, 72, 256, True, None)])
def test_gemm_with_folded_layernorm(M, N, K, f32, act):
    """rowstats + gemm_ln (LayerNorm folded into the GEMM through row statistics and the weight's column sums) against
    F.layer_norm -> F.linear in fp32, on inputs with a large common offset (the term the fold has to cancel)."""
    k = _k()
    g = torch.Generator().manual_seed(M + N)
    x = (torch.randn(M, K, generator=g) * (0.5 + torch.rand(M, 1, generator=g) * 3) + torch.randn(M, 1, generator=g) * 4).to(torch.bfloat16)
    gamma, beta = 1 + 0.3 * torch.randn(K, generator=g), 0.2 * torch.randn(K, generator=g)
    W, b = torch.randn(N, K, generator=g) / K ** 0.5, torch.randn(N, generator=g)
    eps = 1e-6
    ref = torch.nn.functional.linear(torch.nn.functional.layer_norm(x.float(), (K,), gamma, beta, eps), W, b)
    if act == "gelu":
        ref = torch.nn.functional.gelu(ref)
    st = k.rowstats(x.cuda(), eps)
    mu, var = x.float().mean(1), x.float().var(1, unbiased=False)
    assert torch.allclose(st[:, 0].cpu(), mu, atol=1e-4, rtol=1e-4)
    assert torch.allclose(st[:, 1].cpu(), 1 / torch.sqrt(var + eps), rtol=1e-3)
    wf = (W * gamma[None, :]).to(torch.bfloat16)
    out = k.gemm_ln(x.cuda(), wf.cuda(), (W @ beta + b).cuda(), wf.float().sum(1).cuda(), st, act=act,
                    out_dtype=torch.float32 if f32 else torch.bfloat16).cpu().float()
    assert out.shape == ref.shape
    err = (out - ref).abs()
    assert (err <= 2e-2 * ref.abs() + 3e-2).all(), err.max()
    assert rel_l2(out, ref) < 8e-3


@pytest.mark.parametrize("M,C,gamma_on", [(256, 96, True), (1000, 96, True), (37, 192, True), (2048 + 128, 192, False),
                                          (900, 384, True), (256 * 80 + 8, 384, True), (256 * 150, 96, True)])
def test_convnext_mlp_fused(M, C, gamma_on):
    """ConvNeXt block tail in one launch (LN -> pwconv1 -> GELU -> pwconv2 -> gamma -> += into the fp32 stream,
    twin_convnext.py:98-132) against F.layer_norm / F.linear / F.gelu in fp32; ragged M (partial 256-row tiles, a peer CTA
    with no rows), more tiles than CTA pairs (the persistent loop and both weight rings wrap), rows with a large offset."""
    k = _k()
    F = torch.nn.functional
    g = torch.Generator().manual_seed(M + C)
    y = (torch.randn(M, C, generator=g) * (0.5 + torch.rand(M, 1, generator=g) * 2) + torch.randn(M, 1, generator=g) * 2).to(torch.bfloat16)
    lw, lb = 1 + 0.3 * torch.randn(C, generator=g), 0.2 * torch.randn(C, generator=g)
    W1, b1 = torch.randn(4 * C, C, generator=g) / C ** 0.5, 0.5 * torch.randn(4 * C, generator=g)
    W2, b2 = torch.randn(C, 4 * C, generator=g) / (4 * C) ** 0.5, 0.5 * torch.randn(C, generator=g)
    gamma = (0.3 + 0.1 * torch.randn(C, generator=g)) if gamma_on else None
    t0 = torch.randn(M, C, generator=g) * 3
    eps = 1e-6
    h = F.gelu(F.linear(F.layer_norm(y.float(), (C,), lw, lb, eps), W1, b1))
    delta = F.linear(h, W2, b2)
    if gamma is not None:
        delta = delta * gamma
    w1f = (W1 * lw[None, :]).to(torch.bfloat16)
    t = t0.clone().cuda()
    k.convnext_mlp(y.cuda(), w1f.cuda(), w1f.float().sum(1).cuda(), (W1 @ lb + b1).cuda(), W2.to(torch.bfloat16).cuda(), b2.cuda(),
                   None if gamma is None else gamma.cuda(), t, eps)
    got = t.cpu() - t0
    err = (got - delta).abs()
    assert (err <= 2e-2 * delta.abs() + 2e-2).all(), (err.max(), err.argmax())
    assert rel_l2(got, delta) < 8e-3
    # the residual add itself is exact fp32: t - delta_kernel reproduces t0 to rounding
    t2 = t0.clone().cuda()
    k.convnext_mlp(y.cuda(), w1f.cuda(), w1f.float().sum(1).cuda(), (W1 @ lb + b1).cuda(), W2.to(torch.bfloat16).cuda(), b2.cuda(),
                   None if gamma is None else gamma.cuda(), t2, eps)
    assert torch.equal(t, t2)                       # deterministic


@pytest.mark.parametrize("B,C,ho,wo,srcs,relu", [
    (2, 64, 32, 32, [(16, 16), (8, 8), (4, 4)], True),      # Segformer head: 3 coarser levels added to the 1/4 level
    (1, 16, 27, 40, [(12, 20), (7, 9), (27, 40)], False),   # ragged row blocks, non-integer ratios, same-size source
    (1, 8, 6, 6, [(1, 1), (2, 3)], True),                   # degenerate sources (all corners clamp to one pixel)
    (1, 24, 25, 19, [(13, 10), (25, 19)], False),           # ragged, one source already at the output size
    (2, 8, 12, 12, [], True),                               # base only
    (1, 128, 27, 40, [(12, 20), (7, 9), (14, 40)], False),  # 64-channel blocks but non-integer ratios: the general kernel
    (2, 64, 40, 72, [(20, 36), (10, 18), (5, 9)], True),    # staged kernel (exact 2x / 4x / 8x): several tiles per axis, ragged last column tile
    (1, 128, 8, 16, [(4, 8), (2, 4), (1, 2)], False),       # staged, one tile: every window row / column clamps at a map edge
])
def test_resize_sum_affine(B, C, ho, wo, srcs, relu):
    """act((base + sum_k bilinear(src_k)) * scale + shift) against F.interpolate(align_corners=False); with the fusion
    weight applied per level this equals concat + 1x1 fusion conv + BN + ReLU (decode_heads/segformer_head.py:55-64)."""
    k = _k()
    g = torch.Generator().manual_seed(B * C + ho)
    base = torch.randn(B, ho, wo, C, generator=g).to(torch.bfloat16)
    ss = [torch.randn(B, h, w, C, generator=g).to(torch.bfloat16) for h, w in srcs]
    scale = torch.rand(C, generator=g) + 0.5
    shift = torch.randn(C, generator=g)
    ref = base.float()
    for t_ in ss:
        ref = ref + torch.nn.functional.interpolate(t_.float().permute(0, 3, 1, 2), size=(ho, wo), mode="bilinear",
                                                    align_corners=False).permute(0, 2, 3, 1)
    ref = ref * scale + shift
    if relu:
        ref = ref.clamp_min(0)
    out = k.resize_sum_affine(base.cuda(), [t_.cuda() for t_ in ss], srcs, (ho, wo), B, C, scale=scale.cuda(),
                              shift=shift.cuda(), relu=relu).cpu().float().view(B, ho, wo, C)
    assert ((out - ref).abs() <= 0.008 * ref.abs() + 6e-3).all()


def test_normalize_u8_matches_reference_pipeline():
    """HWC uint8 -> normalised fp32 NCHW channels (Normalize_multimodal with norm_by_max + ImageToTensor,
    pipelines/transform.py:2796-2806): RGB with the ImageNet statistics, LiDAR with mean 0 / std 1."""
    k = _k()
    g = torch.Generator().manual_seed(5)
    rgb = torch.randint(0, 256, (2, 24, 36, 3), generator=g, dtype=torch.uint8)
    aux = torch.randint(0, 256, (2, 24, 36, 3), generator=g, dtype=torch.uint8)
    mean, std = [0.485, 0.456, 0.406], [0.229, 0.224, 0.225]
    out = torch.full((2, 6, 24, 36), float("nan"), device="cuda")
    k.normalize_u8(rgb.cuda(), out, 0, mean, std)
    k.normalize_u8(aux.cuda(), out, 3, [0, 0, 0], [1, 1, 1])
    ref = torch.cat((((rgb.float() / 255.0) - torch.tensor(mean)) / torch.tensor(std), aux.float() / 255.0), -1).permute(0, 3, 1, 2)
    assert torch.allclose(out.cpu(), ref, atol=2e-6, rtol=1e-6)


@pytest.mark.parametrize("B,hs,ws,ncls,Ho,Wo,crop", [
    (2, 64, 64, 25, 256, 256, None),        # the bench shape's 4x up-sampling (staged kernel)
    (1, 50, 50, 14, 200, 200, (150, 200)),   # FMB-like: crop of whole_dim_cut
    (1, 17, 23, 19, 40, 61, None),           # non-integer ratios, ragged tiles
    (1, 48, 48, 25, 48, 48, None),           # same size: source window too large to stage -> flat kernel
])
def test_upsample_argmax_vs_torch(B, hs, ws, ncls, Ho, Wo, crop):
    """resize(logits -> image size) + softmax + argmax (+ crop) of encoder_decoder.py:96-117, 329-414 against
    F.interpolate(bilinear, align_corners=False).argmax on the pixels whose top-2 margin is above fp32 noise."""
    k = _k()
    g = torch.Generator().manual_seed(hs * Wo)
    npad = (ncls + 31) // 32 * 32
    lg = torch.randn(B, hs, ws, npad, generator=g)
    up = torch.nn.functional.interpolate(lg[..., :ncls].permute(0, 3, 1, 2), size=(Ho, Wo), mode="bilinear", align_corners=False)
    if crop is not None:
        up = up[:, :, :crop[0], :crop[1]]
    top2 = up.topk(2, dim=1).values
    sure = (top2[:, 0] - top2[:, 1]) > 1e-4
    got = k.upsample_argmax(lg.view(-1, npad).cuda(), B, (hs, ws), ncls, (Ho, Wo), crop).cpu().long()
    assert got.shape == up.argmax(1).shape and int(got.max()) < ncls
    assert torch.equal(got[sure], up.argmax(1)[sure])


@pytest.mark.parametrize("p,C,c_off,H,W", [(4, 3, 0, 32, 48), (4, 3, 3, 32, 48), (16, 3, 0, 64, 32), (2, 3, 1, 8, 12)])
def test_patchify_vs_unfold(p, C, c_off, H, W):
    """NCHW fp32 channels [c_off, c_off+C) -> bf16 rows [(b,py,px), (c,ky,kx)] (PatchEmbed / ConvNeXt stem im2col,
    base/image_encoder.py:662-671, base/twin_convnext.py:295-312): vector kernels (p = 4, 16) and the generic one."""
    k = _k()
    g = torch.Generator().manual_seed(p * 100 + H)
    img = torch.randn(2, 6, H, W, generator=g)
    ref = torch.nn.functional.unfold(img[:, c_off:c_off + C], kernel_size=p, stride=p)          # [B, C*p*p, L]
    ref = ref.transpose(1, 2).reshape(-1, C * p * p).to(torch.bfloat16)
    out = k.patchify(img.cuda(), c_off, C, p).cpu()
    assert torch.equal(out.view(ref.shape), ref)


@pytest.mark.parametrize("K,B,C,grids,act", [
    (7, 2, 96, [(37, 50)], None),                       # ConvNeXt 7x7, ragged tiles, 1.5 channel groups
    (7, 3, 384, [(16, 16)], None),                      # many (channel group, image, tile) items per persistent CTA
    (3, 2, 256, [(32, 32), (16, 16), (8, 8)], "gelu"),  # ConvFFN DWConv: three token grids, shared weights, fused GELU
    (3, 1, 192, [(19, 25)], "relu6"),                   # MobileNetV2 block of the fusion neck
    (3, 2, 256, [(40, 72), (20, 36), (10, 18)], "gelu"),  # 3x3 TMA kernel: several tiles per grid, ragged right / bottom tiles
    (3, 1, 64, [(5, 7)], None),                         # map smaller than one tile
    (3, 2, 40, [(12, 12)], "gelu"),                     # C % 32 != 0: the generic kernel
])
def test_dwconv_vs_torch(K, B, C, grids, act):
    """Depthwise KxK (twin_convnext.py:98-101, adapter_modules_...new.py:456-471, :281-295) against F.conv2d(groups=C)."""
    k = _k()
    g = torch.Generator().manual_seed(K * 100 + C)
    S = sum(h * w for h, w in grids)
    x = torch.randn(B, S, C, generator=g).to(torch.bfloat16)
    wt = torch.randn(C, 1, K, K, generator=g) / K
    bias = torch.randn(C, generator=g)
    w_tap = wt.reshape(C, K * K).t().contiguous()        # [K*K, C] tap-major
    out = k.dwconv(x.cuda(), w_tap.cuda(), bias.cuda(), K, grids, B, C, S * C, S * C, act=act).cpu().float()
    refs, o = [], 0
    for (h, w) in grids:
        xi = x[:, o:o + h * w].float().reshape(B, h, w, C).permute(0, 3, 1, 2)
        y = torch.nn.functional.conv2d(xi, wt, bias, padding=K // 2, groups=C)
        if act == "gelu":
            y = torch.nn.functional.gelu(y)
        elif act == "relu6":
            y = torch.clamp(y, 0, 6)
        refs.append(y.permute(0, 2, 3, 1).reshape(B, h * w, C))
        o += h * w
    ref = torch.cat(refs, 1)
    assert ((out - ref).abs() <= 0.008 * ref.abs() + 4e-3).all(), (out - ref).abs().max().item()


def test_colstats_gate_ln_dual():
    k = _k()
    B, HW, C = 2, 1500, 192
    g = torch.Generator().manual_seed(8)
    o = (torch.randn(B, HW, C, generator=g) + 0.3).to(torch.bfloat16)
    wp = torch.randn(HW, generator=g)
    st = k.colstats(o.reshape(-1, C).cuda(), wp.cuda(), B, HW, C).cpu()
    od = o.double()
    ref = torch.stack((od.sum(1), (od * od).sum(1), (od * wp.double()[None, :, None]).sum(1)), -1)
    assert torch.allclose(st, ref, rtol=1e-5, atol=1e-3)
    a = torch.randn(1000, 2 * C, generator=g).to(torch.bfloat16)
    u = k.gate(a.cuda(), C).float().cpu()
    refu = torch.nn.functional.gelu(a[:, :C].float()) * a[:, C:].float()
    assert (u - refu).abs().max() < 0.03
    x = torch.randn(500, C, generator=g).to(torch.bfloat16)
    w, b = torch.randn(C, generator=g), torch.randn(C, generator=g)
    r = torch.empty_like(x).cuda()
    n = k.layernorm(x.cuda(), w.cuda(), b.cuda(), 1e-5, out2=r).float().cpu()
    refn = torch.nn.functional.layer_norm(x.float(), (C,), w, b, 1e-5)
    assert (n - refn).abs().max() < 0.05 and (r.float().cpu() - (refn + x.float())).abs().max() < 0.06


def test_combine_pool_and_ca_apply():
    k = _k()
    B, H, W, C = 2, 20, 37, 192
    g = torch.Generator().manual_seed(12)
    o = torch.randn(B, H, W, C, generator=g).to(torch.bfloat16)
    lo = torch.randn(B, H, W, C, generator=g).to(torch.bfloat16)
    mu, rstd, gt = torch.randn(B, C, generator=g), torch.rand(B, C, generator=g) + 0.5, torch.rand(B, C, generator=g) + 1
    wp, bp = torch.randn(H * W, generator=g), torch.randn(H * W, generator=g)
    args = (o.reshape(-1, C).cuda(), lo.reshape(-1, C).cuda(), mu.cuda(), rstd.cuda(), gt.cuda(), wp.cuda(), bp.cuda(), 0.7, 1.3,
            B, H, W, C)
    f, ph, pwp = k.combine_pool(*args)
    f2, ph2, pwp2 = k.combine_pool(*args)
    assert torch.equal(ph, ph2) and torch.equal(pwp, pwp2) and torch.equal(f, f2)      # reproducible (no atomics)
    pw = pwp.sum(0)
    ref = 0.7 * (((o.float() - mu[:, None, None]) * rstd[:, None, None]) * wp.view(1, H, W, 1) + bp.view(1, H, W, 1)) \
        * gt[:, None, None] + 1.3 * lo.float()
    fo = f.float().cpu().view(B, H, W, C)
    assert ((fo - ref).abs() <= 0.01 * ref.abs() + 0.02).all()
    assert (ph.cpu() - ref.sum(2)).abs().max() < 0.3 and (pw.cpu() - ref.sum(1)).abs().max() < 0.3
    ah, aw = torch.rand(B, H, C, generator=g), torch.rand(B, W, C, generator=g)
    out = k.ca_apply(f, ah.cuda(), aw.cuda(), B, H, W, C).float().cpu().view(B, H, W, C)
    ref2 = f.float().cpu().view(B, H, W, C) * (1 + aw[:, None] * ah[:, :, None])
    assert ((out - ref2).abs() <= 0.01 * ref2.abs() + 0.02).all()


def test_neck_glue_kernels_vs_torch():
    """The O(B*C^2) steps of the fusion neck (adapter_modules_...new.py:98-107, 250-259, 148-162, 187-215) against
    their torch restatement, from the same partial sums."""
    k = _k()
    g = torch.Generator().manual_seed(31)
    B, ci, heads, HW, nch = 2, 96, 8, 4096, 3
    ch = ci // heads
    # --- gfe_weff: chunk partials -> cosine attention -> softmax -> proj fold
    Sp = torch.randn(nch, B, ci, ci, generator=g) * 30
    nqp, nkp = torch.rand(nch, B, ci, generator=g) * 400 + 50, torch.rand(nch, B, ci, generator=g) * 400 + 50
    temp = torch.rand(heads, generator=g) + 0.5
    wp = torch.randn(ci, ci, generator=g) * 0.1
    s2 = torch.tensor([0.7])
    weff = k.gfe_weff(Sp.cuda(), nqp.cuda(), nkp.cuda(), temp.cuda(), wp.cuda(), s2.cuda(), heads).float().cpu()
    S = Sp.sum(0).view(B, heads, ch, heads, ch).diagonal(dim1=1, dim2=3).permute(0, 3, 1, 2)       # [B, heads, ch, ch]
    qn = nqp.sum(0).sqrt().clamp_min(1e-12).view(B, heads, ch, 1)
    kn = nkp.sum(0).sqrt().clamp_min(1e-12).view(B, heads, 1, ch)
    att = torch.softmax(S / (qn * kn) * temp.view(1, heads, 1, 1), -1)
    ref = (torch.einsum("iha,bhaj->bihj", wp.view(ci, heads, ch), att) * s2).reshape(B, ci, ci)
    assert (weff - ref).abs().max() < 4e-3 * ref.abs().max() + 1e-4
    # --- gffm_softmax
    Ep = torch.randn(nch, B, ci, ci, generator=g) * 3
    ax, ay = k.gffm_softmax(Ep.cuda())
    E = Ep.sum(0)
    assert (ax.float().cpu() - torch.softmax(E, -1)).abs().max() < 4e-3
    assert (ay.float().cpu() - torch.softmax(E.transpose(1, 2), -1)).abs().max() < 4e-3
    # --- ffrm_gate: LayerNorm-over-HW statistics, GAP, 1x1 conv, GroupNorm(32), ReLU, sigmoid
    C = 192
    o = (torch.randn(B, HW, C, generator=g) * 2 + 0.3).to(torch.bfloat16)
    wpix, bpix = torch.randn(HW, generator=g), torch.randn(HW, generator=g)
    wf = torch.randn(C, C, generator=g) * 0.2
    gw, gb = torch.rand(C, generator=g) + 0.5, torch.randn(C, generator=g) * 0.1
    part = k.colstats_part(o.reshape(-1, C).cuda(), wpix.cuda(), B, HW, C)
    mu, rstd, gate = k.ffrm_gate(part, HW, float(wpix.double().sum()), float(bpix.double().mean()), 1e-5, wf.cuda(), gw.cuda(),
                                 gb.cuda(), 32, 1e-5)
    od = o.double()
    mu_r, var_r = od.mean(1), od.var(1, unbiased=False)
    ln = (od - mu_r[:, None]) / torch.sqrt(var_r[:, None] + 1e-5) * wpix.double()[None, :, None] + bpix.double()[None, :, None]
    gap = ln.mean(1).float()
    a = torch.nn.functional.group_norm((gap @ wf.t()).unsqueeze(-1), 32, gw, gb, 1e-5).squeeze(-1)
    gate_r = 1 + torch.sigmoid(torch.relu(a))
    assert torch.allclose(mu.cpu().double(), mu_r, atol=1e-4) and torch.allclose(rstd.cpu().double(), 1 / torch.sqrt(var_r + 1e-5), rtol=1e-4)
    assert (gate.cpu() - gate_r).abs().max() < 2e-3
    # --- ca_vectors
    H, W, mip, ns = 12, 20, 8, 3
    ph = torch.randn(B, H, C, generator=g) * W
    pwp = torch.randn(ns, B, W, C, generator=g) * H / ns
    w1, b1 = torch.randn(mip, C, generator=g) * 0.1, torch.randn(mip, generator=g) * 0.1
    bs, bt = torch.rand(mip, generator=g) + 0.5, torch.randn(mip, generator=g) * 0.1
    wh, bh = torch.randn(C, mip, generator=g) * 0.3, torch.randn(C, generator=g) * 0.1
    ww, bw = torch.randn(C, mip, generator=g) * 0.3, torch.randn(C, generator=g) * 0.1
    ah, aw = k.ca_vectors(ph.cuda(), pwp.cuda(), B, H, W, C, *(t.cuda() for t in (w1, b1, bs, bt, wh, bh, ww, bw)))
    y = torch.cat((ph / W, pwp.sum(0) / H), 1)
    y = (y @ w1.t() + b1) * bs + bt
    y = y * torch.nn.functional.relu6(y + 3) / 6
    assert (ah.cpu() - torch.sigmoid(y[:, :H] @ wh.t() + bh)).abs().max() < 1e-4
    assert (aw.cpu() - torch.sigmoid(y[:, H:] @ ww.t() + bw)).abs().max() < 1e-4


def test_gemm_grouped_per_image_weights():
    """mmsam_gemm_grouped_bf16: rows of image b use weight b (fusion-neck per-image products), one launch."""
    k = _k()
    g = torch.Generator().manual_seed(77)
    G, rows, N, K = 3, 512, 96, 96
    a = torch.randn(G * rows, 3 * K, generator=g).to(torch.bfloat16)        # strided A: columns [2K, 3K) of a wider matrix
    w = (torch.randn(G, N, K, generator=g) * 0.1).to(torch.bfloat16)
    res = torch.randn(G * rows, N, generator=g).to(torch.bfloat16)
    scale = torch.rand(N, generator=g) + 0.5
    out = torch.empty(G * rows, 2 * N, dtype=torch.bfloat16, device="cuda")
    ac = a.cuda()
    k.gemm_grouped(ac[:, 2 * K:], w.cuda(), rows, scale=scale.cuda(), residual=res.cuda(), out=out[:, N:])
    ref = torch.cat([(a[i * rows:(i + 1) * rows, 2 * K:].float() @ w[i].float().t()) for i in range(G)]) * scale + res.float()
    got = out[:, N:].float().cpu()
    assert ((got - ref).abs() <= 0.01 * ref.abs() + 0.02).all()


@pytest.mark.parametrize("C,mode", [(96, "f32"), (384, "f32"), (1024, "f32"), (96, "bf16"), (1024, "bf16"), (192, "patchify")])
def test_layernorm_fp32_stream(C, mode):
    """fp32 input (the fp32 residual streams) -> fp32 or bf16 output, incl. the 2x2 patchify scatter."""
    k = _k()
    g = torch.Generator().manual_seed(C)
    H, W = 6, 10
    x = torch.randn(2 * H * W, C, generator=g) * 3 + 0.5
    w, b = 1 + 0.1 * torch.randn(C, generator=g), 0.1 * torch.randn(C, generator=g)
    ref = torch.nn.functional.layer_norm(x, (C,), w, b, 1e-6)
    if mode == "f32":
        out = k.layernorm(x.cuda(), w.cuda(), b.cuda(), 1e-6, out_dtype=torch.float32)
        assert out.dtype == torch.float32 and (out.cpu() - ref).abs().max() < 1e-4
        xc = x.cuda()
        k.layernorm(xc, w.cuda(), b.cuda(), 1e-6, out=xc)                   # in place
        assert (xc.cpu() - ref).abs().max() < 1e-4
    elif mode == "bf16":
        out = k.layernorm(x.cuda(), w.cuda(), b.cuda(), 1e-6)
        assert out.dtype == torch.bfloat16 and ((out.float().cpu() - ref).abs() <= 0.004 * ref.abs() + 1e-3).all()
    else:
        out = k.layernorm(x.cuda(), w.cuda(), b.cuda(), 1e-6, patchify_hw=(H, W)).float().cpu()
        exp = ref.view(2, H // 2, 2, W // 2, 2, C).permute(0, 1, 3, 2, 4, 5).reshape(2 * (H // 2) * (W // 2), 4 * C)
        assert ((out - exp).abs() <= 0.004 * exp.abs() + 1e-3).all()


@pytest.mark.parametrize("M,N,K,bn", [(1000, 384, 1536, 0), (4096, 1024, 1024, 256), (777, 96, 384, 0), (640, 1024, 512, 128)])
def test_gemm_fp32_residual_stream(M, N, K, bn):
    """fp32 output + fp32 residual, in place (proj / lin2 / pw2 on the fp32 residual streams), with bias and column scale."""
    k = _k()
    g = torch.Generator().manual_seed(M + N)
    a = torch.randn(M, K, generator=g).to(torch.bfloat16)
    w = (torch.randn(N, K, generator=g) / math.sqrt(K)).to(torch.bfloat16)
    bias, scale = torch.randn(N, generator=g), torch.rand(N, generator=g) + 0.5
    x = torch.randn(M, N, generator=g) * 4
    ref = _gemm_ref(a, w, bias=bias, scale=scale, residual=x)
    xc = x.cuda()
    out = k.gemm(a.cuda(), w.cuda(), bias=bias.cuda(), scale=scale.cuda(), residual=xc, out=xc, block_n=bn)
    assert out.dtype == torch.float32 and out.data_ptr() == xc.data_ptr()
    assert (out.cpu().double() - ref).abs().max() < 2e-3 * math.sqrt(K / 64)
    out2 = k.gemm(a.cuda(), w.cuda(), residual=x.cuda(), out_dtype=torch.float32)          # fresh output buffer
    assert (out2.cpu().double() - _gemm_ref(a, w, residual=x)).abs().max() < 2e-3 * math.sqrt(K / 64)


def test_dwconv_fp32_input_and_resize_fp32_source():
    k = _k()
    g = torch.Generator().manual_seed(5)
    B, C, h, w = 2, 96, 20, 37
    x = torch.randn(B, h * w, C, generator=g) * 2
    wt = torch.randn(C, 1, 7, 7, generator=g) / 7
    bias = torch.randn(C, generator=g)
    out = k.dwconv(x.cuda(), wt.reshape(C, 49).t().contiguous().cuda(), bias.cuda(), 7, [(h, w)], B, C, h * w * C, h * w * C)
    assert out.dtype == torch.bfloat16
    xi = x.reshape(B, h, w, C).permute(0, 3, 1, 2)       # fp32 operands, as the reference's conv sees them
    ref = torch.nn.functional.conv2d(xi, wt, bias, padding=3, groups=C).permute(0, 2, 3, 1).reshape(B, h * w, C)
    assert ((out.float().cpu() - ref).abs() <= 0.008 * ref.abs() + 4e-3).all()
    src = torch.randn(B, 8, 8, 128, generator=g)
    base = torch.randn(B, 32, 32, 128, generator=g).to(torch.bfloat16)
    got = k.resize_add_affine(src.cuda(), (8, 8), (32, 32), B, 128, base=base.cuda()).float().cpu()
    up = torch.nn.functional.interpolate(src.permute(0, 3, 1, 2), size=(32, 32), mode="bilinear", align_corners=False).permute(0, 2, 3, 1)
    assert ((got - (up + base.float())).abs() <= 0.008 * (up + base.float()).abs() + 4e-3).all()


@pytest.mark.parametrize("B,C,h,w", [(1, 32, 7, 5), (2, 96, 33, 70), (3, 384, 64, 64), (8, 768, 32, 32), (1, 192, 128, 128),
                                     (2, 96, 16, 32), (1, 64, 17, 33)])
def test_dwconv7_fp32_stream(B, C, h, w):
    """ConvNeXt 7x7 depthwise conv on the fp32 residual stream (twin_convnext.py:98-101): the TMA-staged kernel (halo
    zero-filled by the copy engine, output clipped by the TMA store) against F.conv2d in fp32: maps smaller than one tile,
    ragged right / bottom tiles, both row-group variants (the small-map heuristic picks 8-row tiles), batch strides."""
    k = _k()
    g = torch.Generator().manual_seed(B * 1000 + C + h)
    x = torch.randn(B, h * w, C, generator=g) * 2
    wt = torch.randn(C, 1, 7, 7, generator=g) / 7
    bias = torch.randn(C, generator=g)
    out = k.dwconv(x.cuda(), wt.reshape(C, 49).t().contiguous().cuda(), bias.cuda(), 7, [(h, w)], B, C, h * w * C, h * w * C)
    ref = torch.nn.functional.conv2d(x.reshape(B, h, w, C).permute(0, 3, 1, 2), wt, bias, padding=3, groups=C).permute(0, 2, 3, 1).reshape(B, h * w, C)
    err = (out.float().cpu() - ref).abs()
    assert (err <= 0.004 * ref.abs() + 1e-3).all(), err.max().item()      # bf16 rounding of the output only


@pytest.mark.parametrize("Lq,shapes", [(4096, [(128, 128), (64, 64), (32, 32)]), (21504, [(64, 64)])])
def test_msda_baseline_shapes_vs_oracle(Lq, shapes):
    """Kernel-level parity at the BASELINE shapes: injector (Lq 4096, S 21504, L 3) and extractor
    (Lq 21504, S 4096, L 1), M 16, D 32, P 4, one image: the generic entry point (mmsam_msda_forward, bf16 value + fp32
    locations / weights) and the fused front end (mmsam_msda_fused_bf16) against the CPU oracle."""
    from oracle.msda import ms_deform_attn_core
    k = _k()
    g = torch.Generator().manual_seed(Lq)
    N, M, D, P, L = 1, 16, 32, 4, len(shapes)
    S = sum(h * w for h, w in shapes)
    shp = torch.as_tensor(shapes, dtype=torch.long)
    lsi = torch.cat((shp.new_zeros((1,)), shp.prod(1).cumsum(0)[:-1]))
    value = torch.randn(N, S, M, D, generator=g).to(torch.bfloat16)
    # queries on a grid of cell centres (reference points), offsets of a few pixels, joint softmax over L*P
    side = int(round(math.sqrt(Lq))) if L == 3 else None
    if L == 3:
        ys, xs = torch.meshgrid((torch.arange(side) + 0.5) / side, (torch.arange(side) + 0.5) / side, indexing="ij")
        ref = torch.stack((xs.reshape(-1), ys.reshape(-1)), -1)
    else:
        pts = []
        for hh in (128, 64, 32):
            ys, xs = torch.meshgrid((torch.arange(hh) + 0.5) / hh, (torch.arange(hh) + 0.5) / hh, indexing="ij")
            pts.append(torch.stack((xs.reshape(-1), ys.reshape(-1)), -1))
        ref = torch.cat(pts)
    off = torch.randn(N, Lq, M, L, P, 2, generator=g) * 2.5
    logit = torch.randn(N, Lq, M, L * P, generator=g)
    wh = torch.stack((shp[:, 1], shp[:, 0]), -1).float()
    loc = ref[None, :, None, None, None, :] + off / wh[None, None, None, :, None, :]
    aw = torch.softmax(logit, -1).view(N, Lq, M, L, P)
    want = ms_deform_attn_core(value.float(), shp, loc, aw)
    got = k.msda_forward(value.cuda(), shp.cuda(), lsi.cuda(), loc.cuda(), aw.cuda()).float().cpu()
    assert (got - want).abs().max() < 1e-3 + 0.008 * want.abs().max()           # one bf16 rounding of the output
    qproj = torch.cat((off.reshape(N * Lq, -1), logit.reshape(N * Lq, -1)), 1).contiguous()
    got2 = k.msda_fused(value.view(N, S, M * D).cuda(), shp.cuda(), lsi.cuda(), qproj.cuda(), ref.contiguous().cuda(), M, L, P).float().cpu()
    assert (got2 - want).abs().max() < 1e-3 + 0.008 * want.abs().max()
