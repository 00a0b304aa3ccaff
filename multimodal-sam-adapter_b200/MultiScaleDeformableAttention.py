"""Drop-in for the reference's compiled extension `MultiScaleDeformableAttention`
(segmentation/ops/src/vision.cpp:13-16), forwarding to the C ABI (mmsam_msda_forward / mmsam_msda_backward).

    import MultiScaleDeformableAttention as MSDA
    out = MSDA.ms_deform_attn_forward(value, spatial_shapes, level_start_index, sampling_loc, attn_weight, im2col_step)

Conventions kept from ops/src/cuda/ms_deform_attn_cuda.cu:28-52: all tensors contiguous CUDA tensors
(else RuntimeError), batch % min(batch, im2col_step) == 0, value/loc/weight share a dtype (the autograd
Function casts), a new tensor is returned, the launch goes on the current stream without a host sync.
im2col_step is accepted for signature compatibility; the whole batch is one launch.
"""
import torch

from . import kernels as _K


def ms_deform_attn_forward(value, spatial_shapes, level_start_index, sampling_loc, attn_weight, im2col_step):
    for name, t in (("value", value), ("spatial_shapes", spatial_shapes), ("level_start_index", level_start_index),
                    ("sampling_loc", sampling_loc), ("attn_weight", attn_weight)):
        if not t.is_contiguous():
            raise RuntimeError(f"{name} tensor has to be contiguous")
        if not t.is_cuda:
            raise RuntimeError("Not implemented on the CPU" if name == "value" else f"{name} must be a CUDA tensor")
    batch = value.shape[0]
    step = min(batch, int(im2col_step)) if batch > 0 else 1
    if step <= 0 or batch % step != 0:
        raise RuntimeError("batch(%d) must divide im2col_step(%d)" % (batch, step))
    try:
        return _K.msda_forward(value, spatial_shapes, level_start_index, sampling_loc, attn_weight)
    except _K._lib.MMSamError as e:
        raise RuntimeError(str(e)) from e


def ms_deform_attn_backward(value, spatial_shapes, level_start_index, sampling_loc, attn_weight, grad_output,
                            im2col_step):
    """-> [grad_value, grad_sampling_loc, grad_attn_weight] (ops/src/cuda/ms_deform_attn_cuda.cu:83-153); fp32 / fp64."""
    for name, t in (("value", value), ("spatial_shapes", spatial_shapes), ("level_start_index", level_start_index),
                    ("sampling_loc", sampling_loc), ("attn_weight", attn_weight), ("grad_output", grad_output)):
        if not t.is_contiguous():
            raise RuntimeError(f"{name} tensor has to be contiguous")
        if not t.is_cuda:
            raise RuntimeError("Not implemented on the CPU" if name == "value" else f"{name} must be a CUDA tensor")
    batch = value.shape[0]
    step = min(batch, int(im2col_step)) if batch > 0 else 1
    if step <= 0 or batch % step != 0:
        raise RuntimeError("batch(%d) must divide im2col_step(%d)" % (batch, step))
    try:
        return list(_K.msda_backward(value, spatial_shapes, level_start_index, sampling_loc, attn_weight, grad_output))
    except _K._lib.MMSamError as e:
        raise RuntimeError(str(e)) from e
