"""Thin torch-tensor front ends over the C ABI (include/mmsam_b200.h).

torch is used for device memory and the current CUDA stream only; every op below launches one of
this repo's sm_100a kernels and raises if the tensors are not on a CUDA device.
"""
import os

import torch

from . import _lib

_DT = {torch.float32: _lib.F32, torch.float16: _lib.F16, torch.bfloat16: _lib.BF16, torch.float64: _lib.F64}

LAUNCHES = 0  # number of kernel launches issued through this module (bench.py reports it)


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _ptr(t):
    return 0 if t is None else t.data_ptr()


def _need_cuda(*ts):
    """Every tensor must live on the CURRENT CUDA device: the kernels launch on its current stream and their
    shared-memory attributes are set per device (run under `with torch.cuda.device(t.device)`; the engine does)."""
    cur = None
    for t in ts:
        if t is None:
            continue
        if not t.is_cuda:
            raise _lib.MMSamError("mmsam_b200 kernels need CUDA tensors (there is no CPU fallback)")
        if cur is None:
            cur = torch.cuda.current_device()
        if t.device.index != cur:
            raise _lib.MMSamError(f"tensor on cuda:{t.device.index} but the current device is cuda:{cur}: wrap the call in "
                                  "`with torch.cuda.device(tensor.device)`")


def _count(n=1):
    global LAUNCHES
    LAUNCHES += n


def msda_forward(value, spatial_shapes, level_start_index, sampling_loc, attn_weight, out=None):
    """value [N,S,M,D]; spatial_shapes [L,2] i64 (device); level_start_index [L] i64 (device);
    sampling_loc [N,Lq,M,L,P,2]; attn_weight [N,Lq,M,L,P]  ->  [N,Lq,M*D] in value's dtype."""
    _need_cuda(value, spatial_shapes, level_start_index, sampling_loc, attn_weight)
    for t in (value, spatial_shapes, level_start_index, sampling_loc, attn_weight):
        if not t.is_contiguous():
            raise _lib.MMSamError("msda_forward: all tensors must be contiguous")
    if spatial_shapes.dtype != torch.int64 or level_start_index.dtype != torch.int64:
        raise _lib.MMSamError("msda_forward: spatial_shapes / level_start_index must be int64")
    if sampling_loc.dtype != attn_weight.dtype:
        raise _lib.MMSamError("msda_forward: sampling_loc and attn_weight must share a dtype")
    N, S, M, D = value.shape
    _, Lq, _, L, P, _ = sampling_loc.shape
    if out is None:
        out = torch.empty((N, Lq, M * D), dtype=value.dtype, device=value.device)
    rc = _lib.load().mmsam_msda_forward(
        _ptr(value), _ptr(spatial_shapes), _ptr(level_start_index), _ptr(sampling_loc), _ptr(attn_weight),
        _ptr(out), N, S, M, D, Lq, L, P, _DT[value.dtype], _DT[sampling_loc.dtype], _stream())
    _lib.check(rc, "mmsam_msda_forward")
    _count()
    return out


def msda_backward(value, spatial_shapes, level_start_index, sampling_loc, attn_weight, grad_output):
    """Gradients of msda_forward (fp32 / fp64, all floating tensors in one dtype):
    -> (grad_value [N,S,M,D], grad_sampling_loc [N,Lq,M,L,P,2], grad_attn_weight [N,Lq,M,L,P])."""
    _need_cuda(value, spatial_shapes, level_start_index, sampling_loc, attn_weight, grad_output)
    for t in (value, spatial_shapes, level_start_index, sampling_loc, attn_weight, grad_output):
        if not t.is_contiguous():
            raise _lib.MMSamError("msda_backward: all tensors must be contiguous")
    if spatial_shapes.dtype != torch.int64 or level_start_index.dtype != torch.int64:
        raise _lib.MMSamError("msda_backward: spatial_shapes / level_start_index must be int64")
    if value.dtype not in (torch.float32, torch.float64) or any(t.dtype != value.dtype for t in (sampling_loc, attn_weight, grad_output)):
        raise _lib.MMSamError("msda_backward: value / sampling_loc / attn_weight / grad_output must all be fp32 or all fp64")
    N, S, M, D = value.shape
    _, Lq, _, L, P, _ = sampling_loc.shape
    gv = torch.empty_like(value)
    gl = torch.empty_like(sampling_loc)
    ga = torch.empty_like(attn_weight)
    rc = _lib.load().mmsam_msda_backward(
        _ptr(value), _ptr(spatial_shapes), _ptr(level_start_index), _ptr(sampling_loc), _ptr(attn_weight), _ptr(grad_output),
        _ptr(gv), _ptr(gl), _ptr(ga), N, S, M, D, Lq, L, P, _DT[value.dtype], _stream())
    _lib.check(rc, "mmsam_msda_backward")
    _count()
    return gv, gl, ga


def layernorm(x, gamma, beta, eps, out=None, row_map=None, out_rows=None, patchify_hw=None, out2=None,
              out_dtype=torch.bfloat16):
    """x bf16 or fp32 [..., C] (rows contiguous) -> LN over C (bf16, or fp32 for an fp32 input when out_dtype says so).
    row_map (int32 [rows]) scatters rows into an `out` of out_rows rows (rows never written keep their previous
    contents). patchify_hw=(H, W): rows are (b,y,x) and the result is the 2x2-patchified [rows/4, 4C] matrix."""
    _need_cuda(x, gamma, beta)
    C = x.shape[-1]
    x2 = x.reshape(-1, C)
    rows = x2.shape[0]
    ps_h = ps_w = 0
    if patchify_hw is not None:
        ps_h, ps_w = patchify_hw
    if out is None:
        if patchify_hw is not None:
            out = torch.empty((rows // 4, 4 * C), dtype=out_dtype, device=x.device)
        elif row_map is None:
            out = torch.empty(x2.shape, dtype=out_dtype, device=x.device)
        else:
            out = torch.zeros((out_rows, C), dtype=out_dtype, device=x.device)
    rc = _lib.load().mmsam_layernorm(
        _ptr(x2), _DT[x2.dtype], _ptr(gamma), _ptr(beta), _ptr(out), _DT[out.dtype], _ptr(out2), _ptr(row_map), rows, C,
        x2.stride(0), out.stride(0), float(eps), ps_h, ps_w, _stream())
    _lib.check(rc, "mmsam_layernorm")
    _count()
    if row_map is not None or patchify_hw is not None:
        return out
    return out.view(x.shape)


ACT = {None: 0, "none": 0, "gelu": 1, "relu": 2, "relu6": 3}


def gemm(a, w, bias=None, act=None, scale=None, residual=None, out=None, out_dtype=torch.bfloat16,
         row_map=None, out_rows=None, pixel_shuffle=None, block_n=0, max_ctas=0):
    """out = epilogue(a[M,K] @ w[N,K]^T); a, w bf16 with contiguous K. See include/mmsam_b200.h."""
    _need_cuda(a, w, bias, scale, residual, out, row_map)
    M, K = a.shape
    N = w.shape[0]
    assert w.shape[1] == K and a.stride(1) == 1 and w.stride(1) == 1
    row_mode, ps_h, ps_w, ps_c = 0, 0, 0, 0
    n_out, m_out = N, M
    if row_map is not None:
        row_mode, m_out = 1, (out_rows if out_rows is not None else M)
    if pixel_shuffle is not None:
        ps_h, ps_w = pixel_shuffle
        ps_c = N // 4
        row_mode, m_out, n_out = 2, M * 4, ps_c
    if out is None:
        out = torch.empty((m_out, n_out), dtype=out_dtype, device=a.device)
    rc = _lib.load().mmsam_gemm_bf16(
        _ptr(a), a.stride(0), _ptr(w), w.stride(0), _ptr(bias), _ptr(scale), _ptr(residual),
        1 if (residual is not None and residual.dtype == torch.float32) else 0,
        residual.stride(0) if residual is not None else 0, _ptr(out), out.stride(0), M, N, K, ACT[act],
        1 if out.dtype == torch.float32 else 0, row_mode, _ptr(row_map), ps_h, ps_w, ps_c, block_n, max_ctas,
        _stream())
    _lib.check(rc, "mmsam_gemm_bf16")
    _count()
    return out


def gemm_grouped(a, w, rows_per_group, bias=None, scale=None, residual=None, out=None, act=None, block_n=0, max_ctas=0):
    """Per-group weights in one launch: rows [g*rows_per_group, (g+1)*rows_per_group) of a [M, K] use w[g] of w [G, N, K]
    (bf16, contiguous); rows_per_group % 256 == 0. See include/mmsam_b200.h (mmsam_gemm_grouped_bf16)."""
    _need_cuda(a, w, bias, scale, residual, out)
    M, K = a.shape
    G, N, K2 = w.shape
    assert K2 == K and w.is_contiguous() and a.stride(1) == 1 and G * rows_per_group == M
    if out is None:
        out = torch.empty((M, N), dtype=torch.bfloat16, device=a.device)
    rc = _lib.load().mmsam_gemm_grouped_bf16(
        _ptr(a), a.stride(0), _ptr(w), K, _ptr(bias), _ptr(scale), _ptr(residual),
        residual.stride(0) if residual is not None else 0, _ptr(out), out.stride(0), M, N, K, rows_per_group, ACT[act],
        block_n, max_ctas, _stream())
    _lib.check(rc, "mmsam_gemm_grouped_bf16")
    _count()
    return out


def rowstats(x, eps, out=None):
    """x bf16 [rows, C] (rows contiguous) -> fp32 [rows, 2] = (mean, rstd) of every row (LayerNorm statistics)."""
    _need_cuda(x)
    rows, C = x.shape
    if out is None:
        out = torch.empty((rows, 2), dtype=torch.float32, device=x.device)
    rc = _lib.load().mmsam_rowstats_bf16(_ptr(x), _ptr(out), rows, C, x.stride(0), float(eps), _stream())
    _lib.check(rc, "mmsam_rowstats_bf16")
    _count()
    return out


def gemm_ln(a, w, bias, colsum, rowstat, act=None, residual=None, out=None, out_dtype=torch.bfloat16, block_n=0,
            max_ctas=0):
    """LayerNorm folded into the GEMM: out = act(rstd * (a @ w^T - mean * colsum) + bias) (+ residual);
    w = bf16(gamma * W), colsum = w.sum(1), bias = W @ beta + b (see include/mmsam_b200.h)."""
    _need_cuda(a, w, bias, colsum, rowstat, residual, out)
    M, K = a.shape
    N = w.shape[0]
    assert w.shape[1] == K and a.stride(1) == 1 and w.stride(1) == 1 and rowstat.shape[0] == M
    if out is None:
        out = torch.empty((M, N), dtype=out_dtype, device=a.device)
    rc = _lib.load().mmsam_gemm_ln_bf16(
        _ptr(a), a.stride(0), _ptr(w), w.stride(0), _ptr(bias), _ptr(colsum), _ptr(rowstat), _ptr(residual),
        residual.stride(0) if residual is not None else 0, _ptr(out), out.stride(0), M, N, K, ACT[act],
        1 if out.dtype == torch.float32 else 0, block_n, max_ctas, _stream())
    _lib.check(rc, "mmsam_gemm_ln_bf16")
    _count()
    return out


def convnext_mlp(y, w1, colsum1, bias1, w2, bias2, gamma, t, eps, max_ctas=0):
    """ConvNeXt block tail in one launch: t += gamma * (GELU(LN(y) @ w1^T + b1) @ w2^T + b2), in place on the fp32 stream t;
    the LayerNorm affine is folded into w1 / colsum1 / bias1 (see include/mmsam_b200.h)."""
    _need_cuda(y, w1, colsum1, bias1, w2, bias2, gamma, t)
    M, C = y.shape
    assert t.shape == (M, C) and t.dtype == torch.float32 and y.dtype == torch.bfloat16 and y.stride(1) == 1 and t.stride(1) == 1
    assert w1.shape == (4 * C, C) and w2.shape == (C, 4 * C) and w1.is_contiguous() and w2.is_contiguous()
    rc = _lib.load().mmsam_convnext_mlp_bf16(_ptr(y), y.stride(0), _ptr(w1), _ptr(colsum1), _ptr(bias1), _ptr(w2), _ptr(bias2),
                                             _ptr(gamma), _ptr(t), t.stride(0), M, C, float(eps), max_ctas, _stream())
    _lib.check(rc, "mmsam_convnext_mlp_bf16")
    _count()
    return t


def relpos_table(rel_pos, size):
    """SAM get_rel_pos for q_size == k_size == size (base/image_encoder.py:554-584): linear
    interpolation of the [L, 64] table to 2*size-1 rows when L differs; row r = q - k + size - 1.
    Returns a zero-padded bf16 [pad16(2*size-1), 64] table for attention()."""
    n = 2 * size - 1
    t = rel_pos.detach().float()
    if t.shape[0] != n:
        t = torch.nn.functional.interpolate(t.t()[None], size=n, mode="linear")[0].t()
    pad = (n + 15) // 16 * 16
    out = torch.zeros((pad, t.shape[1]), dtype=torch.bfloat16, device=rel_pos.device)
    out[:n] = t.to(torch.bfloat16)
    return out.contiguous()


def attention(qkv, num_heads, hw, tab_h=None, tab_w=None, out=None, max_ctas=0, out_map=None, out_rows=None):
    """qkv bf16 [Bp, T, 3*nh*64] (qkv Linear output) -> [Bp, T, nh*64]; hw = (Kh, Kw), T = Kh*Kw.
    out_map (int32 [Bp*T]) scatters the output rows into a [out_rows, nh*64] matrix (-1 = drop)."""
    _need_cuda(qkv, tab_h, tab_w, out_map)
    Bp, T, C3 = qkv.shape
    hd = C3 // (3 * num_heads)
    if hd != 64 or not qkv.is_contiguous():
        raise _lib.MMSamError("attention: head_dim must be 64 and qkv contiguous")
    if out is None:
        if out_map is not None:
            out = torch.empty((out_rows, num_heads * hd), dtype=torch.bfloat16, device=qkv.device)
        else:
            out = torch.empty((Bp, T, num_heads * hd), dtype=torch.bfloat16, device=qkv.device)
    rc = _lib.load().mmsam_attention_bf16(_ptr(qkv), _ptr(out), _ptr(out_map), _ptr(tab_h), _ptr(tab_w), Bp, T,
                                          num_heads, hw[0], hw[1], float(hd) ** -0.5, max_ctas, _stream())
    _lib.check(rc, "mmsam_attention_bf16")
    _count()
    return out


def attention_window(qkv, num_heads, B, H, W, tab_h=None, tab_w=None, out=None, max_ctas=0):
    """SAM window attention (14 x 14 windows) with window_unpartition fused into the store: qkv bf16
    [B * nwh * nww, 196, 3 * nh * 64] in window order -> the token map [B * H * W, nh * 64] (base/image_encoder.py:399-416,
    529-551)."""
    _need_cuda(qkv, tab_h, tab_w)
    nwh, nww = (H + 13) // 14, (W + 13) // 14
    Bp, T, C3 = qkv.shape
    if T != 196 or Bp != B * nwh * nww or C3 != 3 * num_heads * 64 or not qkv.is_contiguous():
        raise _lib.MMSamError("attention_window: qkv must be contiguous [B * ceil(H/14) * ceil(W/14), 196, 3 * heads * 64]")
    if out is None:
        out = torch.empty((B * H * W, num_heads * 64), dtype=torch.bfloat16, device=qkv.device)
    rc = _lib.load().mmsam_attention_window_bf16(_ptr(qkv), _ptr(out), _ptr(tab_h), _ptr(tab_w), B, H, W, num_heads, 64.0 ** -0.5,
                                                 max_ctas, _stream())
    _lib.check(rc, "mmsam_attention_window_bf16")
    _count()
    return out


class MsdaGeometry:
    """Host-side geometry for the shared-memory staged MSDeformAttn kernel (mmsam_msda_fused_staged_bf16): level
    shapes, query grids (row-major, reference points = cell centres, adapter_modules_...new.py:397-431), the anchor
    grid / tile that partitions the normalised plane into CTA regions, and the per-level / per-head prior of the
    sampling offsets taken from the sampling_offsets bias [M, L, P, 2] (ops/modules/ms_deform_attn.py:64-74)."""

    def __init__(self, level_shapes, qgrid_shapes, anchor, tile, offset_bias, n_heads, n_levels, n_points=4, margin=2):
        import ctypes
        L, M = n_levels, n_heads
        assert len(level_shapes) == L
        b = offset_bias.detach().float().cpu().view(M, L, n_points, 2)
        pmin = b.min(2)[0].permute(1, 0, 2).contiguous()                    # [L, M, 2]
        ext = (b.max(2)[0] - b.min(2)[0]).max(0)[0].contiguous()            # [L, 2]: max over heads
        self.levels = (ctypes.c_int * (2 * L))(*[int(v) for hw in level_shapes for v in hw])
        self.qgrids = (ctypes.c_int * (2 * len(qgrid_shapes)))(*[int(v) for hw in qgrid_shapes for v in hw])
        self.n_qgrids = len(qgrid_shapes)
        self.pmin = (ctypes.c_float * (L * M * 2))(*pmin.view(-1).tolist())
        self.ext = (ctypes.c_float * (L * 2))(*ext.view(-1).tolist())
        self.anchor, self.tile, self.margin = (int(anchor[0]), int(anchor[1])), (int(tile[0]), int(tile[1])), int(margin)
        self.S = sum(h * w for h, w in level_shapes)
        self.Lq = sum(h * w for h, w in qgrid_shapes)
        self.unsupported = False


def msda_fused(value, spatial_shapes, level_start_index, qproj, ref_xy, n_heads, n_levels, n_points=4, out=None,
               geom=None):
    """value bf16 [N,S,M*D]; qproj fp32 [N*Lq, >= M*L*P*3] (offsets | logits); ref_xy fp32 [Lq,2].
    geom (MsdaGeometry, optional): use the shared-memory staged kernel when the configuration supports it."""
    import ctypes
    _need_cuda(value, spatial_shapes, level_start_index, qproj, ref_xy)
    N, S, MD = value.shape
    D = MD // n_heads
    Lq = ref_xy.shape[0]
    assert qproj.shape[0] == N * Lq and qproj.dtype == torch.float32 and qproj.stride(1) == 1
    if out is None:
        out = torch.empty((N, Lq, MD), dtype=value.dtype, device=value.device)
    if geom is not None and not geom.unsupported and os.environ.get("MMSAM_MSDA_STAGED", "1") != "0":
        assert geom.S == S and geom.Lq == Lq
        rc = _lib.load().mmsam_msda_fused_staged_bf16(
            _ptr(value), ctypes.cast(geom.levels, ctypes.c_void_p), _ptr(qproj), qproj.stride(0), _ptr(ref_xy), _ptr(out),
            N, S, n_heads, D, Lq, n_levels, n_points, geom.n_qgrids, ctypes.cast(geom.qgrids, ctypes.c_void_p),
            geom.anchor[0], geom.anchor[1], geom.tile[0], geom.tile[1], ctypes.cast(geom.pmin, ctypes.c_void_p),
            ctypes.cast(geom.ext, ctypes.c_void_p), geom.margin, _stream())
        if rc != -3:                              # MMSAM_ERR_UNSUPPORTED: fall through to the L1-gather kernel
            _lib.check(rc, "mmsam_msda_fused_staged_bf16")
            _count()
            return out
        geom.unsupported = True
    rc = _lib.load().mmsam_msda_fused_bf16(
        _ptr(value), _ptr(spatial_shapes), _ptr(level_start_index), _ptr(qproj), qproj.stride(0), _ptr(ref_xy),
        _ptr(out), N, S, n_heads, D, Lq, n_levels, n_points, _stream())
    _lib.check(rc, "mmsam_msda_fused_bf16")
    _count()
    return out


def dwconv(x, w_tap_major, bias, ksize, grids, B, C, in_bstride, out_bstride, act=None, out=None,
           in_offs=None, out_offs=None):
    """Depthwise conv over channels-last bf16 maps. grids: [(H, W), ...] (<= 3) inside each batch item."""
    import ctypes
    _need_cuda(x, w_tap_major, bias)
    if out is None:
        out = torch.empty(x.shape, dtype=torch.bfloat16, device=x.device)
    n = len(grids)
    hw = (ctypes.c_int * (2 * n))(*[v for g in grids for v in g])
    if in_offs is None:
        in_offs, o = [], 0
        for (h, w) in grids:
            in_offs.append(o)
            o += h * w * C
    if out_offs is None:
        out_offs = in_offs
    io = (ctypes.c_longlong * n)(*in_offs)
    oo = (ctypes.c_longlong * n)(*out_offs)
    rc = _lib.load().mmsam_dwconv(_ptr(x), _DT[x.dtype], _ptr(out), _ptr(w_tap_major), _ptr(bias), B, C, ksize, n,
                                       ctypes.cast(hw, ctypes.c_void_p), ctypes.cast(io, ctypes.c_void_p),
                                       ctypes.cast(oo, ctypes.c_void_p), in_bstride, out_bstride, ACT[act], _stream())
    _lib.check(rc, "mmsam_dwconv")
    _count()
    return out


def normalize_u8(img_u8, out, c_off, mean, std, prescale=1.0 / 255.0):
    """img_u8 uint8 [B, H, W, C] (HWC, CUDA) -> out[:, c_off:c_off+C] of the fp32 NCHW input: (v * prescale - mean) / std."""
    import ctypes
    _need_cuda(img_u8, out)
    assert img_u8.dtype == torch.uint8 and img_u8.is_contiguous() and out.dtype == torch.float32 and out.is_contiguous()
    B, H, W, C = img_u8.shape
    assert out.shape[0] == B and tuple(out.shape[2:]) == (H, W)
    m = (ctypes.c_float * C)(*[float(v) for v in mean])
    sd = (ctypes.c_float * C)(*[float(v) for v in std])
    rc = _lib.load().mmsam_normalize_u8(_ptr(img_u8), _ptr(out), B, H, W, C, out.shape[1], c_off, ctypes.cast(m, ctypes.c_void_p),
                                        ctypes.cast(sd, ctypes.c_void_p), float(prescale), _stream())
    _lib.check(rc, "mmsam_normalize_u8")
    _count()
    return out


def patchify(img, c_off, C, p, out=None):
    """img fp32 NCHW [B, Ctot, H, W] -> bf16 [B*(H/p)*(W/p), C*p*p]."""
    _need_cuda(img)
    assert img.dtype == torch.float32 and img.is_contiguous()
    B, Ctot, H, W = img.shape
    if out is None:
        out = torch.empty((B * (H // p) * (W // p), C * p * p), dtype=torch.bfloat16, device=img.device)
    rc = _lib.load().mmsam_patchify_f32(_ptr(img), _ptr(out), B, Ctot, c_off, C, H, W, p, _stream())
    _lib.check(rc, "mmsam_patchify_f32")
    _count()
    return out


class U8Modality:
    """One modality's decoded frames as they leave the loader: uint8 HWC [B, Hs, Ws, C] on the device, plus the
    normalisation of Normalize_multimodal (pipelines/transform.py:2717-2815): (v * prescale - mean) / std, prescale =
    1/255 with norm_by_max; to_rgb reverses the channel order first; pad_val fills the rows / columns Pad_multimodal adds."""

    def __init__(self, data, mean, std, prescale=1.0 / 255.0, to_rgb=False, pad_val=0.0):
        self.data, self.mean, self.std = data, [float(v) for v in mean], [float(v) for v in std]
        self.prescale, self.to_rgb, self.pad_val = float(prescale), bool(to_rgb), float(pad_val)


def patchify_u8(mod, H, W, p, out=None):
    """U8Modality -> bf16 [B*(H/p)*(W/p), C*p*p] patch rows of the H x W (padded) normalised image."""
    import ctypes
    img = mod.data
    _need_cuda(img)
    assert img.dtype == torch.uint8 and img.is_contiguous() and img.dim() == 4
    B, Hs, Ws, C = img.shape
    if out is None:
        out = torch.empty((B * (H // p) * (W // p), C * p * p), dtype=torch.bfloat16, device=img.device)
    m = (ctypes.c_float * C)(*mod.mean[:C])
    sd = (ctypes.c_float * C)(*mod.std[:C])
    rc = _lib.load().mmsam_patchify_u8(_ptr(img), _ptr(out), B, Hs, Ws, C, H, W, p, ctypes.cast(m, ctypes.c_void_p),
                                       ctypes.cast(sd, ctypes.c_void_p), mod.prescale, mod.pad_val, 1 if mod.to_rgb else 0, _stream())
    _lib.check(rc, "mmsam_patchify_u8")
    _count()
    return out


def resize_logits(logits, B, hw, out_hw):
    """fp32 channels-last logits [B*h*w, ld] -> [B*Ho*Wo, ld], bilinear, align_corners=False."""
    _need_cuda(logits)
    ld = logits.shape[-1]
    out = torch.empty((B * out_hw[0] * out_hw[1], ld), dtype=torch.float32, device=logits.device)
    rc = _lib.load().mmsam_resize_logits_f32(_ptr(logits), _ptr(out), B, hw[0], hw[1], ld, out_hw[0], out_hw[1], _stream())
    _lib.check(rc, "mmsam_resize_logits_f32")
    _count()
    return out


def slide_merge(crop_logits, B, hw, ncls, frame_hw, boxes, want_preds=False):
    """crop_logits fp32 [ncrops*B*h*w, ld] (crop j of image b at index j*B+b) + boxes [(y1, x1, y2, x2)] -> uint8 labels
    [B, H, W], or the averaged fp32 logits [B*H*W, ld] when want_preds (a resize to ori_shape follows)."""
    import ctypes
    _need_cuda(crop_logits)
    ld = crop_logits.shape[-1]
    H, W = frame_hw
    bx = (ctypes.c_int * (4 * len(boxes)))(*[int(v) for b in boxes for v in b])
    labels = preds = None
    if want_preds:
        preds = torch.empty((B * H * W, ld), dtype=torch.float32, device=crop_logits.device)
    else:
        labels = torch.empty((B, H, W), dtype=torch.uint8, device=crop_logits.device)
    rc = _lib.load().mmsam_slide_merge_f32(_ptr(crop_logits), _ptr(labels), _ptr(preds), B, hw[0], hw[1], ld, ncls, H, W,
                                           len(boxes), ctypes.cast(bx, ctypes.c_void_p), _stream())
    _lib.check(rc, "mmsam_slide_merge_f32")
    _count()
    return preds if want_preds else labels


def resize_add_affine(src, src_hw, out_hw, B, C, base=None, scale=None, shift=None, out=None,
                      src_bstride=None, lds=None, base_bstride=None, ldb=None, out_bstride=None, ldo=None):
    """Channels-last: out = (base + bilinear(src)) * scale + shift (src bf16 or fp32; base / out bf16). Strides in elements."""
    _need_cuda(src, base, scale, shift, out)
    Hs, Ws = src_hw
    Ho, Wo = out_hw
    lds = C if lds is None else lds
    ldb = C if ldb is None else ldb
    ldo = C if ldo is None else ldo
    src_bstride = Hs * Ws * lds if src_bstride is None else src_bstride
    base_bstride = Ho * Wo * ldb if base_bstride is None else base_bstride
    out_bstride = Ho * Wo * ldo if out_bstride is None else out_bstride
    if out is None:
        out = torch.empty((B, Ho, Wo, C), dtype=torch.bfloat16, device=src.device)
    rc = _lib.load().mmsam_resize_add_affine(
        _ptr(src), _DT[src.dtype], _ptr(base), _ptr(scale), _ptr(shift), _ptr(out), B, Hs, Ws, Ho, Wo, C, src_bstride, lds,
        base_bstride, ldb, out_bstride, ldo, _stream())
    _lib.check(rc, "mmsam_resize_add_affine")
    _count()
    return out


def resize_sum_affine(base, srcs, src_hws, out_hw, B, C, scale=None, shift=None, relu=False, out=None):
    """Dense channels-last bf16: out = act((base + sum_k bilinear(srcs[k])) * scale + shift); <= 3 sources."""
    import ctypes
    _need_cuda(base, scale, shift, out, *srcs)
    Ho, Wo = out_hw
    assert len(srcs) == len(src_hws) <= 3
    hw = (ctypes.c_int * max(2 * len(srcs), 2))(*[int(v) for s in src_hws for v in s])
    ptrs = [_ptr(s) for s in srcs] + [0] * (3 - len(srcs))
    if out is None:
        out = torch.empty((B * Ho * Wo, C), dtype=torch.bfloat16, device=(base if base is not None else srcs[0]).device)
    rc = _lib.load().mmsam_resize_sum_affine_bf16(_ptr(base), len(srcs), ptrs[0], ptrs[1], ptrs[2],
                                                  ctypes.cast(hw, ctypes.c_void_p), _ptr(scale), _ptr(shift),
                                                  1 if relu else 0, _ptr(out), B, Ho, Wo, C, _stream())
    _lib.check(rc, "mmsam_resize_sum_affine_bf16")
    _count()
    return out


def upsample_argmax(logits, B, hw, ncls, out_hw, crop_hw=None, out=None):
    """logits fp32 [B*h*w, ldl] -> uint8 labels [B, Hc, Wc] (bilinear to out_hw, argmax, crop)."""
    _need_cuda(logits)
    hs, ws = hw
    Ho, Wo = out_hw
    Hc, Wc = crop_hw if crop_hw is not None else out_hw
    if out is None:
        out = torch.empty((B, Hc, Wc), dtype=torch.uint8, device=logits.device)
    rc = _lib.load().mmsam_upsample_argmax_f32(_ptr(logits), _ptr(out), B, hs, ws, logits.stride(0), ncls, Ho, Wo,
                                               Hc, Wc, _stream())
    _lib.check(rc, "mmsam_upsample_argmax_f32")
    _count()
    return out


def confusion(pred, gt, ncls, ignore_index=255, out=None):
    """uint8 pred / gt -> int64 [ncls, ncls] confusion counts (rows = gt), accumulated into `out`."""
    _need_cuda(pred, gt)
    assert pred.dtype == torch.uint8 and gt.dtype == torch.uint8 and pred.numel() == gt.numel()
    if out is None:
        out = torch.zeros((ncls, ncls), dtype=torch.int64, device=pred.device)
    rc = _lib.load().mmsam_confusion_u8(_ptr(pred.contiguous()), _ptr(gt.contiguous()), _ptr(out), pred.numel(), ncls,
                                        ignore_index, _stream())
    _lib.check(rc, "mmsam_confusion_u8")
    _count()
    return out


def pack_conv3x3_weight(w, groups):
    """Conv2d weight [Cout, Cin/groups, 3, 3] -> the block layout mmsam_conv3x3_bf16 expects (bf16, on w's device)."""
    Cout, cgi = w.shape[0], w.shape[1]
    Cin = cgi * groups
    cgo = Cout // groups
    KC = _lib.load().mmsam_conv3x3_kblocks(Cin, Cout, groups)
    NS = _lib.load().mmsam_conv3x3_nstride(Cin, Cout, groups)      # output channels per n-tile (<= 64)
    nn = (Cout + NS - 1) // NS
    wf = w.detach().float().cpu().reshape(Cout, cgi, 9)
    out = torch.zeros((nn, 9, KC, 64, 64), dtype=torch.float32)
    co = torch.arange(Cout)
    nt, r = co // NS, co % NS
    kwin = (((nt * NS) // cgo) * cgi) // 8 * 8      # first input channel of the tile's window (16-byte aligned)
    base = (co // cgo) * cgi - kwin                 # window column of each output channel's group start
    for ci in range(cgi):
        col = base + ci
        out[nt, :, col // 64, r, col % 64] = wf[:, ci, :]
    return out.reshape(nn * 9 * KC * 64, 64).to(torch.bfloat16).to(w.device).contiguous()


def conv3x3(x, w_packed, B, H, W, Cin, Cout, groups, out=None, max_ctas=0):
    """x bf16 [B*H*W, Cin] channels-last -> [B*H*W, Cout]; grouped 3x3, pad 1, no bias."""
    _need_cuda(x, w_packed)
    if out is None:
        out = torch.empty((B * H * W, Cout), dtype=torch.bfloat16, device=x.device)
    rc = _lib.load().mmsam_conv3x3_bf16(_ptr(x), _ptr(w_packed), _ptr(out), B, H, W, Cin, Cout, groups, max_ctas, _stream())
    _lib.check(rc, "mmsam_conv3x3_bf16")
    _count()
    return out


def gram(x, ld, qoff, koff, n, B, HW, blk=0, norms=False):
    """Per-pixel-chunk partial Gram sums: S_part fp32 [nchunks, B, n, n] (and nq_part, nk_part fp32 [nchunks, B, n] when
    norms). With blk > 0 only the block-diagonal elements are written. The chunks are added in a fixed order by the
    consumers (gfe_weff / gffm_softmax): no atomics, bit-reproducible."""
    _need_cuda(x)
    lib = _lib.load()
    nch = lib.mmsam_gram_chunks(n, B, HW, 1 if norms else 0)
    S = torch.empty((nch, B, n, n), dtype=torch.float32, device=x.device)
    nq = nk = None
    if norms:
        nq = torch.empty((nch, B, n), dtype=torch.float32, device=x.device)
        nk = torch.empty((nch, B, n), dtype=torch.float32, device=x.device)
    rc = lib.mmsam_gram_bf16(_ptr(x), ld, qoff, koff, n, B, HW, blk, _ptr(S), _ptr(nq), _ptr(nk), _stream())
    _lib.check(rc, "mmsam_gram_bf16")
    _count()
    return (S, nq, nk) if norms else S


def gfe_weff(S_part, nq_part, nk_part, temperature, wproj, scale2, heads):
    """AttentionBase softmax + proj fold (see include/mmsam_b200.h) -> weff bf16 [B, ci, ci]."""
    _need_cuda(S_part, nq_part, nk_part, temperature, wproj, scale2)
    nch, B, ci, _ = S_part.shape
    weff = torch.empty((B, ci, ci), dtype=torch.bfloat16, device=S_part.device)
    rc = _lib.load().mmsam_gfe_weff_bf16(_ptr(S_part), _ptr(nq_part), _ptr(nk_part), nch, B, ci, heads, _ptr(temperature),
                                         _ptr(wproj), _ptr(scale2), _ptr(weff), _stream())
    _lib.check(rc, "mmsam_gfe_weff_bf16")
    _count()
    return weff


def gffm_softmax(E_part):
    """-> (ax, ay) bf16 [B, ci, ci]: row softmax of E and of E^T."""
    _need_cuda(E_part)
    nch, B, ci, _ = E_part.shape
    ax = torch.empty((B, ci, ci), dtype=torch.bfloat16, device=E_part.device)
    ay = torch.empty_like(ax)
    rc = _lib.load().mmsam_gffm_softmax_bf16(_ptr(E_part), nch, B, ci, _ptr(ax), _ptr(ay), _stream())
    _lib.check(rc, "mmsam_gffm_softmax_bf16")
    _count()
    return ax, ay


def colstats_part(o, wpix, B, HW, C):
    """-> fp32 [nchunks, B, C, 3] per-chunk partials {sum o, sum o^2, sum o*w[pix]}."""
    _need_cuda(o, wpix)
    lib = _lib.load()
    if wpix.numel() != HW:
        raise _lib.MMSamError(f"colstats: the LayerNorm-over-HW weight has {wpix.numel()} entries, the map {HW} pixels")
    nch = lib.mmsam_colstats_chunks(HW)
    part = torch.empty((nch, B, C, 3), dtype=torch.float32, device=o.device)
    rc = lib.mmsam_colstats_bf16(_ptr(o), _ptr(wpix), _ptr(part), B, HW, C, _stream())
    _lib.check(rc, "mmsam_colstats_bf16")
    _count()
    return part


def ffrm_gate(part, HW, sum_w, mean_b, ln_eps, wffrm, gn_w, gn_b, groups, gn_eps):
    """colstats partials -> (mu, rstd, gate) fp32 [B, C] (see include/mmsam_b200.h: mmsam_ffrm_gate_f32)."""
    _need_cuda(part, wffrm, gn_w, gn_b)
    nch, B, C, _ = part.shape
    mu = torch.empty((B, C), dtype=torch.float32, device=part.device)
    rstd = torch.empty_like(mu)
    gate_v = torch.empty_like(mu)
    rc = _lib.load().mmsam_ffrm_gate_f32(_ptr(part), nch, B, HW, C, float(sum_w), float(mean_b), float(ln_eps), _ptr(wffrm),
                                         _ptr(gn_w), _ptr(gn_b), groups, float(gn_eps), _ptr(mu), _ptr(rstd), _ptr(gate_v),
                                         _stream())
    _lib.check(rc, "mmsam_ffrm_gate_f32")
    _count()
    return mu, rstd, gate_v


def ca_vectors(ph, pw_part, B, H, W, C, w1, b1, bn_s, bn_t, wh, bh, ww, bw):
    """Pooled sums -> coordinate-attention vectors ah fp32 [B, H, C], aw fp32 [B, W, C]."""
    _need_cuda(ph, pw_part, w1, b1, bn_s, bn_t, wh, bh, ww, bw)
    ah = torch.empty((B, H, C), dtype=torch.float32, device=ph.device)
    aw = torch.empty((B, W, C), dtype=torch.float32, device=ph.device)
    rc = _lib.load().mmsam_ca_vectors_f32(_ptr(ph), _ptr(pw_part), pw_part.shape[0], B, H, W, C, w1.shape[0], _ptr(w1), _ptr(b1),
                                          _ptr(bn_s), _ptr(bn_t), _ptr(wh), _ptr(bh), _ptr(ww), _ptr(bw), _ptr(ah), _ptr(aw),
                                          _stream())
    _lib.check(rc, "mmsam_ca_vectors_f32")
    _count()
    return ah, aw


def colstats(o, wpix, B, HW, C):
    """-> fp64 [B, C, 3] = {sum o, sum o^2, sum o*w[pix]} over the HW pixels of each image (chunks added in order)."""
    return colstats_part(o, wpix, B, HW, C).double().sum(0)


def gate(a, C, out=None):
    _need_cuda(a)
    rows = a.shape[0]
    if out is None:
        out = torch.empty((rows, C), dtype=torch.bfloat16, device=a.device)
    rc = _lib.load().mmsam_gate_bf16(_ptr(a), _ptr(out), rows, C, _stream())
    _lib.check(rc, "mmsam_gate_bf16")
    _count()
    return out


def combine_pool(o, lo, mu, rstd, gate_v, wpix, bpix, s1, s2, B, H, W, C):
    """-> f bf16 [B*H*W, C], ph fp32 [B,H,C] (row sums), pw_part fp32 [strips, B, W, C] (per-strip column sums)."""
    _need_cuda(o, lo, mu, rstd, gate_v, wpix, bpix)
    if wpix.numel() != H * W or bpix.numel() != H * W:
        raise _lib.MMSamError(f"combine_pool: LayerNorm-over-HW affine has {wpix.numel()} entries, the map {H * W} pixels")
    lib = _lib.load()
    RS = lib.mmsam_combine_pool_rows(H)
    ns = (H + RS - 1) // RS
    f = torch.empty((B * H * W, C), dtype=torch.bfloat16, device=o.device)
    ph = torch.empty((B, H, C), dtype=torch.float32, device=o.device)
    pwp = torch.empty((ns, B, W, C), dtype=torch.float32, device=o.device)
    rc = lib.mmsam_combine_pool_bf16(_ptr(o), _ptr(lo), _ptr(mu), _ptr(rstd), _ptr(gate_v), _ptr(wpix), _ptr(bpix),
                                     float(s1), float(s2), _ptr(f), _ptr(ph), _ptr(pwp), B, H, W, C, _stream())
    _lib.check(rc, "mmsam_combine_pool_bf16")
    _count()
    return f, ph, pwp


def ca_apply(f, ah, aw, B, H, W, C, out=None):
    _need_cuda(f, ah, aw)
    if out is None:
        out = torch.empty_like(f)
    rc = _lib.load().mmsam_ca_apply_bf16(_ptr(f), _ptr(ah), _ptr(aw), _ptr(out), B, H, W, C, _stream())
    _lib.check(rc, "mmsam_ca_apply_bf16")
    _count()
    return out
