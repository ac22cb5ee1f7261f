"""`BilinearSample` with the reference's module signature (networks/backbone.py:453-475), backed by
the hand-written gather kernel instead of 4 elementwise kernels + stack + F.grid_sample."""
import torch
import torch.nn as nn
from torch.autograd import Function

from . import ops


class _BilinearSampleFunction(Function):
    @staticmethod
    def forward(ctx, grid_feat, grid_coord, scale_rate, point_major_out, order=None):
        ctx.scale_rate = scale_rate
        ctx.hw = (grid_feat.shape[2], grid_feat.shape[3])
        ctx.save_for_backward(grid_coord)
        return ops.bilinear_gather_forward(grid_feat, grid_coord, scale_rate, point_major_out, order)

    @staticmethod
    def backward(ctx, grad_out):
        (grid_coord,) = ctx.saved_tensors
        grad_grid = None
        if ctx.needs_input_grad[0]:
            grad_grid = ops.bilinear_gather_backward(grad_out, grid_coord, ctx.scale_rate, ctx.hw[0], ctx.hw[1])
        return grad_grid, None, None, None, None


class BilinearSample(nn.Module):
    """forward(grid_feat (BS, C, H, W), grid_coord (BS, N, 2, S)) -> pc_feat (BS, C, N, S).

    `point_major_out=True` returns the same values in channels_last strides (each point's C features
    contiguous) — the layout VoxelMaxPool consumes fastest; shapes and values are unchanged."""

    point_major_out = False

    def __init__(self, in_dim, scale_rate):
        super(BilinearSample, self).__init__()
        self.scale_rate = scale_rate

    def forward(self, grid_feat, grid_coord, order=None):
        """Reference signature plus an optional `order` (an ops.PoolPlan of the same coordinates, e.g. the one the
        neighbouring VoxelMaxPool uses): the points are then visited in cell order — same values, faster."""
        return _BilinearSampleFunction.apply(grid_feat.float(), grid_coord, tuple(self.scale_rate),
                                             bool(self.point_major_out), order)
