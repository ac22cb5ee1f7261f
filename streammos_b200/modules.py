"""`MSDeformAttn` with the reference's module interface (deformattn/modules/ms_deform_attn.py:30-116): same constructor,
parameter names (checkpoints load unchanged), initialisation and forward signature.

What changes is the forward between the three linear projections: the reference runs softmax, a stack, a division, a
broadcast add (five elementwise kernels and their (N, Lq, M, L, P, 2) intermediates) and then the sampling core. Here, for
inference on CUDA tensors (no autograd), ONE kernel takes the raw sampling offsets, the attention logits and the reference
points and does all of it (ops.ms_deform_attn_fused_forward). With autograd enabled the reference's own
sequence runs, ending in MSDeformAttnFunction (our forward/backward kernels)."""
import math
import warnings

import torch
import torch.nn.functional as F
from torch import nn
from torch.nn.init import constant_, xavier_uniform_

from . import ops
from .functions import MSDeformAttnFunction


def _is_power_of_2(n):
    if (not isinstance(n, int)) or (n < 0):
        raise ValueError("invalid input for _is_power_of_2: {} (type: {})".format(n, type(n)))
    return (n & (n - 1) == 0) and n != 0


def msdeformattn_forward(self, query, reference_points, input_flatten, input_spatial_shapes, input_level_start_index,
                         input_padding_mask=None):
    """forward of the reference module (ms_deform_attn.py:78-116); `self` is a reference or a streammos_b200 MSDeformAttn.
    query (N, Lq, C), reference_points (N, Lq, L, 2 | 4), input_flatten (N, sum H_l W_l, C) -> (N, Lq, C)."""
    N, Len_q, _ = query.shape
    N, Len_in, _ = input_flatten.shape
    fused = (query.is_cuda and not torch.is_grad_enabled() and query.dtype in (torch.float32, torch.float64) and
             reference_points.shape[-1] in (2, 4))
    if not fused:  # the module asserts on the device here (a host sync); the fused path leaves shape errors to the kernel
        assert (input_spatial_shapes[:, 0] * input_spatial_shapes[:, 1]).sum() == Len_in
    value = self.value_proj(input_flatten)
    if input_padding_mask is not None:
        value = value.masked_fill(input_padding_mask[..., None], float(0))
        query = query.masked_fill(input_padding_mask[..., None], float(0))
    value = value.view(N, Len_in, self.n_heads, self.d_model // self.n_heads)
    sampling_offsets = self.sampling_offsets(query).view(N, Len_q, self.n_heads, self.n_levels, self.n_points, 2)
    attention_weights = self.attention_weights(query).view(N, Len_q, self.n_heads, self.n_levels * self.n_points)
    if fused:
        output = ops.ms_deform_attn_fused_forward(value.contiguous(), input_spatial_shapes.contiguous(),
                                                  input_level_start_index.contiguous(), sampling_offsets.contiguous(),
                                                  attention_weights.contiguous(),
                                                  reference_points.to(value.dtype).contiguous())
        return self.output_proj(output)
    attention_weights = F.softmax(attention_weights, -1).view(N, Len_q, self.n_heads, self.n_levels, self.n_points)
    if reference_points.shape[-1] == 2:
        offset_normalizer = torch.stack([input_spatial_shapes[..., 1], input_spatial_shapes[..., 0]], -1)
        sampling_locations = reference_points[:, :, None, :, None, :] \
            + sampling_offsets / offset_normalizer[None, None, None, :, None, :]
    elif reference_points.shape[-1] == 4:
        sampling_locations = reference_points[:, :, None, :, None, :2] \
            + sampling_offsets / self.n_points * reference_points[:, :, None, :, None, 2:] * 0.5
    else:
        raise ValueError(
            'Last dim of reference_points must be 2 or 4, but get {} instead.'.format(reference_points.shape[-1]))
    output = MSDeformAttnFunction.apply(value, input_spatial_shapes, input_level_start_index, sampling_locations,
                                        attention_weights, self.im2col_step)
    return self.output_proj(output)


class MSDeformAttn(nn.Module):
    def __init__(self, d_model=256, n_levels=4, n_heads=8, n_points=4):
        super().__init__()
        if d_model % n_heads != 0:
            raise ValueError('d_model must be divisible by n_heads, but got {} and {}'.format(d_model, n_heads))
        if not _is_power_of_2(d_model // n_heads):
            warnings.warn("You'd better set d_model in MSDeformAttn to make the dimension of each attention head a power of 2 "
                          "which is more efficient in our CUDA implementation.")
        self.im2col_step = 256
        self.d_model, self.n_levels, self.n_heads, self.n_points = d_model, n_levels, n_heads, n_points
        self.sampling_offsets = nn.Linear(d_model, n_heads * n_levels * n_points * 2)
        self.attention_weights = nn.Linear(d_model, n_heads * n_levels * n_points)
        self.value_proj = nn.Linear(d_model, d_model)
        self.output_proj = nn.Linear(d_model, d_model)
        self._reset_parameters()

    def _reset_parameters(self):  # ms_deform_attn.py:62-76
        constant_(self.sampling_offsets.weight.data, 0.)
        thetas = torch.arange(self.n_heads, dtype=torch.float32) * (2.0 * math.pi / self.n_heads)
        grid_init = torch.stack([thetas.cos(), thetas.sin()], -1)
        grid_init = (grid_init / grid_init.abs().max(-1, keepdim=True)[0]).view(self.n_heads, 1, 1, 2) \
            .repeat(1, self.n_levels, self.n_points, 1)
        for i in range(self.n_points):
            grid_init[:, :, i, :] *= i + 1
        with torch.no_grad():
            self.sampling_offsets.bias = nn.Parameter(grid_init.view(-1))
        constant_(self.attention_weights.weight.data, 0.)
        constant_(self.attention_weights.bias.data, 0.)
        xavier_uniform_(self.value_proj.weight.data)
        constant_(self.value_proj.bias.data, 0.)
        xavier_uniform_(self.output_proj.weight.data)
        constant_(self.output_proj.bias.data, 0.)

    forward = msdeformattn_forward
