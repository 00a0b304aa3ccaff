"""Mirror of segmentation/ops/functions/ms_deform_attn_func.py:19-50 (the autograd Function boundary)."""
import torch
from torch.autograd import Function

from ... import MultiScaleDeformableAttention as MSDA


class MSDeformAttnFunction(Function):
    @staticmethod
    def forward(ctx, value, value_spatial_shapes, value_level_start_index, sampling_locations, attention_weights,
                im2col_step):
        # the reference casts the aux tensors to value's dtype before the call (func.py:26-27)
        attention_weights = attention_weights.type_as(value).contiguous()
        sampling_locations = sampling_locations.type_as(value).contiguous()
        ctx.im2col_step = im2col_step
        ctx.save_for_backward(value, value_spatial_shapes, value_level_start_index, sampling_locations, attention_weights)
        return MSDA.ms_deform_attn_forward(value, value_spatial_shapes, value_level_start_index,
                                           sampling_locations, attention_weights, im2col_step)

    @staticmethod
    def backward(ctx, grad_output):
        # func.py:39-50: gradients for value / sampling_locations / attention_weights, None for the index tensors and the step
        value, shapes, lsi, loc, attw = ctx.saved_tensors
        gv, gl, ga = MSDA.ms_deform_attn_backward(value, shapes, lsi, loc, attw, grad_output.contiguous(), ctx.im2col_step)
        return gv, None, None, gl, ga, None
