"""Mirror of segmentation/ops/functions/ms_deform_attn_func.py:19-50 (the autograd Function boundary)."""
import torch
from torch.autograd import Function

from ... import MultiScaleDeformableAttention as MSDA


class MSDeformAttnFunction(Function):
    @staticmethod
    def forward(ctx, value, value_spatial_shapes, value_level_start_index, sampling_locations, attention_weights,
                im2col_step):
        # the reference casts the aux tensors to value's dtype before the call (func.py:26-27)
        attention_weights = attention_weights.type_as(value)
        sampling_locations = sampling_locations.type_as(value)
        return MSDA.ms_deform_attn_forward(value, value_spatial_shapes, value_level_start_index,
                                           sampling_locations.contiguous(), attention_weights.contiguous(), im2col_step)

    @staticmethod
    def backward(ctx, grad_output):
        raise RuntimeError("MSDeformAttnFunction.backward: inference-only build")
