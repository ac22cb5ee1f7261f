"""Drop-in for the compiled module `MultiScaleDeformableAttention` (deformattn/src/vision.cpp:13-16,
deformattn/src/ms_deform_attn.h:20-61): the two functions deformattn/functions/ms_deform_attn_func.py
imports as `MSDA`."""
from . import ops


def _step(batch, im2col_step):
    step = min(int(batch), int(im2col_step))
    # ms_deform_attn_cuda.cu:50-52
    assert step > 0 and batch % step == 0, "batch(%d) must divide im2col_step(%d)" % (batch, step)


def ms_deform_attn_forward(value, spatial_shapes, level_start_index, sampling_loc, attn_weight, im2col_step):
    """-> (B, Lq, M*D). `im2col_step` only chunked the reference's launches; one launch covers the batch."""
    _step(value.size(0), im2col_step)
    return ops.ms_deform_attn_forward(value, spatial_shapes, level_start_index, sampling_loc, attn_weight)


def ms_deform_attn_backward(value, spatial_shapes, level_start_index, sampling_loc, attn_weight, grad_output,
                            im2col_step):
    """-> [grad_value, grad_sampling_loc, grad_attn_weight]"""
    _step(value.size(0), im2col_step)
    return list(ops.ms_deform_attn_backward(value, spatial_shapes, level_start_index, sampling_loc, attn_weight,
                                            grad_output))
