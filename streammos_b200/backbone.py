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
    contiguous) — the layout VoxelMaxPool consumes fastest; shapes and values are unchanged.

    `auto_order`: with the reference signature (no `order=`) the points are visited in the cell order of the pooling
    plan of (grid_coord, this grid's (H, W), scale_rate), taken from the plan cache — in the reference's cascade every
    gather samples the grid a pooling call with exactly these coordinates, size and scale wrote or will write
    (multi_view_encoder.py:395-417, StreamMOS.py:105), so the plan exists already or is built here for the pool that
    follows. Same values, fewer cache lines per warp load."""

    point_major_out = False
    auto_order = True

    def __init__(self, in_dim, scale_rate):
        super(BilinearSample, self).__init__()
        self.scale_rate = scale_rate

    def forward(self, grid_feat, grid_coord, order=None):
        """Reference signature plus an optional `order` (an ops.PoolPlan of the same coordinates, e.g. the one the
        neighbouring VoxelMaxPool uses): the points are then visited in cell order — same values, faster."""
        if order is None and self.auto_order and self.point_major_out and grid_coord.is_cuda and \
                grid_coord.dim() == 4 and grid_coord.size(3) == 1 and grid_coord.size(2) == 2 and \
                grid_coord.dtype == torch.float32:
            order = ops.cached_pool_plan(grid_coord, tuple(grid_feat.shape[2:]), tuple(self.scale_rate))
        return _BilinearSampleFunction.apply(grid_feat.float(), grid_coord, tuple(self.scale_rate),
                                             bool(self.point_major_out), order)


def _bn_affine(bn):
    """Eval-mode BatchNorm2d as torch applies it: y = x * alpha + beta."""
    invstd = torch.rsqrt(bn.running_var.float() + bn.eps)
    w = bn.weight.float() if bn.affine else torch.ones_like(invstd)
    b = bn.bias.float() if bn.affine else torch.zeros_like(invstd)
    alpha = invstd * w
    return alpha, b - bn.running_mean.float() * alpha


class PointNet(nn.Module):
    """networks/backbone.py:199-231, same sub-module layout (checkpoints load unchanged)."""

    def __init__(self, cin, cout, pre_bn=False, post_act=True):
        super(PointNet, self).__init__()
        layers = [nn.BatchNorm2d(cin)] if pre_bn else []
        layers += [nn.Conv2d(cin, cout, kernel_size=1, stride=1, padding=0, dilation=1, bias=False), nn.BatchNorm2d(cout)]
        if post_act:
            layers.append(nn.ReLU(inplace=True))
        self.layer = nn.Sequential(*layers)

    def forward(self, x):
        return self.layer(x)


class PointNetStacker(nn.Module):
    """networks/backbone.py:233-250 with the reference's constructor, parameter names and training behaviour.
    In eval mode on a CUDA tensor the configuration StreamMOS builds (models/StreamMOS.py:77: cin -> 64 -> 64,
    pre_bn=True, stack_num=2, post_act=True) runs as ONE fused kernel (smos_point_stem_forward: layer 2 on the tcgen05
    tensor cores as a 3xTF32 split, within 1e-5 of the fp32 layers) instead of seven;
    every other case (training: batch statistics and autograd; other shapes) runs the torch layers, as the
    reference does."""

    point_major_out = False  # fused path: (B, C, N, 1) result with channels_last strides (dropin.install sets it)

    def __init__(self, cin, cout, pre_bn=False, post_act=True, stack_num=1):
        super(PointNetStacker, self).__init__()
        if stack_num == 1:
            layers = [PointNet(cin=cin, cout=cout, pre_bn=pre_bn, post_act=post_act)]
        else:
            layers = [PointNet(cin=cin, cout=cout, pre_bn=pre_bn, post_act=True)]
            for _ in range(1, stack_num - 1):
                layers.append(PointNet(cin=cout, cout=cout, pre_bn=False, post_act=True))
            layers.append(PointNet(cin=cout, cout=cout, pre_bn=False, post_act=post_act))
        self.layer = nn.Sequential(*layers)
        self._fusable = (stack_num == 2 and post_act and cout == 64 and cin <= 16)
        self._pre_bn = pre_bn

    def fused_parameters(self):
        """(bn0, w1, bn1, w2, bn2) for ops.point_stem_forward. The BatchNorm affines are cached and recomputed when any
        parameter or buffer changed (in-place update, load_state_dict, .to()): a dozen tiny torch kernels otherwise
        cost as much as the fused kernel itself."""
        l0, l1 = self.layer[0].layer, self.layer[1].layer
        k = 1 if self._pre_bn else 0
        tensors = list(self.parameters()) + list(self.buffers())
        key = tuple((t.data_ptr(), t._version) for t in tensors)
        if getattr(self, "_fused_key", None) != key:
            with torch.no_grad():
                bn0 = _bn_affine(l0[0]) if self._pre_bn else None
                self._fused = (bn0, l0[k].weight.detach().reshape(l0[k].weight.shape[0], -1).contiguous(),
                               _bn_affine(l0[k + 1]),
                               l1[0].weight.detach().reshape(l1[0].weight.shape[0], -1).contiguous(), _bn_affine(l1[1]))
            self._fused_key = key
        return self._fused

    def forward(self, x):
        if self._fusable and not self.training and x.is_cuda and x.dtype == torch.float32 and x.dim() == 4 and \
                x.size(3) == 1 and not torch.is_grad_enabled():
            return ops.point_stem_forward(x, *self.fused_parameters(), point_major_out=bool(self.point_major_out))
        return self.layer(x)
