"""ctypes binding of libmmsam_b200.so (the C ABI declared in include/mmsam_b200.h).

There is no CPU or library fallback: if the shared library is missing every kernel call raises.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libmmsam_b200.so")

F32, F16, BF16, F64 = 0, 1, 2, 3

_c = ctypes
_vp, _i, _ll, _f = _c.c_void_p, _c.c_int, _c.c_longlong, _c.c_float

# name -> argtypes; must list every symbol include/mmsam_b200.h declares
SIGNATURES = {
    "mmsam_arch": [],
    "mmsam_msda_forward": [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _i, _vp],
    "mmsam_msda_backward": [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _vp],
    "mmsam_layernorm": [_vp, _i, _vp, _vp, _vp, _i, _vp, _vp, _ll, _i, _ll, _ll, _f, _i, _i, _vp],
    "mmsam_gemm_bf16": [_vp, _ll, _vp, _ll, _vp, _vp, _vp, _i, _ll, _vp, _ll, _i, _i, _i, _i, _i, _i, _vp,
                        _i, _i, _i, _i, _i, _vp],
    "mmsam_gemm_grouped_bf16": [_vp, _ll, _vp, _ll, _vp, _vp, _vp, _ll, _vp, _ll, _i, _i, _i, _i, _i, _i, _i, _vp],
    "mmsam_rowstats_bf16": [_vp, _vp, _ll, _i, _ll, _f, _vp],
    "mmsam_gemm_ln_bf16": [_vp, _ll, _vp, _ll, _vp, _vp, _vp, _vp, _ll, _vp, _ll, _i, _i, _i, _i, _i, _i, _i, _vp],
    "mmsam_convnext_mlp_bf16": [_vp, _ll, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _ll, _i, _i, _f, _i, _vp],
    "mmsam_msda_fused_bf16": [_vp, _vp, _vp, _vp, _ll, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _vp],
    "mmsam_msda_fused_staged_bf16": [_vp, _vp, _vp, _ll, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _vp, _i, _i, _i, _i,
                                     _vp, _vp, _i, _vp],
    "mmsam_dwconv": [_vp, _i, _vp, _vp, _vp, _i, _i, _i, _i, _vp, _vp, _vp, _ll, _ll, _i, _vp],
    "mmsam_normalize_u8": [_vp, _vp, _i, _i, _i, _i, _i, _i, _vp, _vp, _f, _vp],
    "mmsam_patchify_u8": [_vp, _vp, _i, _i, _i, _i, _i, _i, _i, _vp, _vp, _f, _f, _i, _vp],
    "mmsam_resize_logits_f32": [_vp, _vp, _i, _i, _i, _i, _i, _i, _vp],
    "mmsam_slide_merge_f32": [_vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _vp, _vp],
    "mmsam_patchify_f32": [_vp, _vp, _i, _i, _i, _i, _i, _i, _i, _vp],
    "mmsam_resize_add_affine": [_vp, _i, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _ll, _ll, _ll, _ll, _ll, _ll, _vp],
    "mmsam_resize_sum_affine_bf16": [_vp, _i, _vp, _vp, _vp, _vp, _vp, _vp, _i, _vp, _i, _i, _i, _i, _vp],
    "mmsam_upsample_argmax_f32": [_vp, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _i, _vp],
    "mmsam_confusion_u8": [_vp, _vp, _vp, _ll, _i, _i, _vp],
    "mmsam_conv3x3_kblocks": [_i, _i, _i],
    "mmsam_conv3x3_nstride": [_i, _i, _i],
    "mmsam_conv3x3_bf16": [_vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _vp],
    "mmsam_gram_bf16": [_vp, _ll, _i, _i, _i, _i, _i, _i, _vp, _vp, _vp, _vp],
    "mmsam_gram_chunks": [_i, _i, _i, _i],
    "mmsam_gfe_weff_bf16": [_vp, _vp, _vp, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp],
    "mmsam_gffm_softmax_bf16": [_vp, _i, _i, _i, _vp, _vp, _vp],
    "mmsam_ffrm_gate_f32": [_vp, _i, _i, _i, _i, _c.c_double, _c.c_double, _f, _vp, _vp, _vp, _i, _f, _vp, _vp, _vp, _vp],
    "mmsam_ca_vectors_f32": [_vp, _vp, _i, _i, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp],
    "mmsam_colstats_chunks": [_i],
    "mmsam_colstats_bf16": [_vp, _vp, _vp, _i, _i, _i, _vp],
    "mmsam_gate_bf16": [_vp, _vp, _ll, _i, _vp],
    "mmsam_combine_pool_rows": [_i],
    "mmsam_combine_pool_bf16": [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _f, _f, _vp, _vp, _vp, _i, _i, _i, _i, _vp],
    "mmsam_ca_apply_bf16": [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp],
    "mmsam_attention_bf16": [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _f, _i, _vp],
    "mmsam_attention_window_bf16": [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _f, _i, _vp],
}

_lib = None


class MMSamError(RuntimeError):
    pass


def load():
    """Load the shared library (building nothing: use build.py / __graft_entry__.build())."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise MMSamError(
                f"{LIB_PATH} not found: the CUDA extension is required (no CPU fallback). "
                "Run `python __graft_entry__.py` or `python multimodal-sam-adapter_b200/build.py`.")
        lib = ctypes.CDLL(LIB_PATH)
        for name, argtypes in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.argtypes = argtypes
            fn.restype = _c.c_int
        _lib = lib
    return _lib


_ERR = {-1: "bad argument", -2: "unsupported dtype", -3: "unsupported configuration", -4: "CUDA driver entry point unavailable"}


def check(rc, what):
    if rc != 0:
        if rc < 0:
            raise MMSamError(f"{what}: {_ERR.get(rc, 'error')} ({rc})")
        raise MMSamError(f"{what}: CUDA error {rc}")
