"""Drop-in for the reference's `deep_point` package (deep_point/__init__.py:15-65).

Same names, argument meaning and assertions: `VoxelMaxPool(pcds_feat, pcds_ind, output_size,
scale_rate)` and `VoxelMaxPoolFunction`. Differences that do not change results: the output
is produced by one output-stationary kernel (no zeros/full fills, no metadata uploads), the
backward re-uses the forward's pooling plan instead of `voxel_max_idx`, plans are shared between
calls that pass the same coordinate tensor (plan_cache.py), and CPU tensors raise
(the reference silently ran its serial C++ loop on them, deep_point/__init__.py:38-40).

  pcds_feat  (BS, C, N, 1)      pcds_ind (BS, N, D=2, 1)
  voxel_out  (BS, C, H, W)
"""
import torch
from torch.autograd import Function

from . import ops


class VoxelMaxPoolFunction(Function):
    @staticmethod
    def forward(ctx, pcds_feat, pcds_ind, output_size, scale_rate, plan=None):
        assert pcds_feat.dtype == pcds_ind.dtype
        assert pcds_feat.dim() == 4
        assert pcds_ind.dim() == 4
        assert pcds_feat.size(2) == pcds_ind.size(1)
        assert pcds_ind.size(2) == len(output_size)
        assert pcds_ind.size(2) == len(scale_rate)
        if not pcds_feat.is_cuda:
            raise RuntimeError("deep_point.VoxelMaxPool: CPU tensors are not supported by the B200 build "
                               "(no CPU fallback); the CPU restatement lives in oracle/ for tests only")
        if plan is None:
            # reference signature: the plan comes from the cache shared with the other pools / gathers of the scan
            # that pass the same coordinate tensor (plan_cache.py); built here on a miss
            plan = ops.cached_pool_plan(pcds_ind, output_size, scale_rate)
        else:
            assert (plan.B, plan.N, plan.H, plan.W) == (pcds_ind.size(0), pcds_ind.size(1), output_size[0],
                                                        output_size[1]), "plan does not match this call"
        voxel_out = ops.voxel_maxpool_forward(pcds_feat, plan)
        ctx.plan = plan
        ctx.input_shape = pcds_feat.shape
        ctx.save_for_backward(pcds_feat, voxel_out)
        return voxel_out

    @staticmethod
    def backward(ctx, grad_voxel_out):
        pcds_feat, voxel_out = ctx.saved_tensors
        if ctx.needs_input_grad[0]:
            grad_pcds_feat = ops.voxel_maxpool_backward(pcds_feat, ctx.plan, voxel_out, grad_voxel_out.contiguous())
            return grad_pcds_feat, None, None, None, None
        return None, None, None, None, None


def VoxelMaxPool(pcds_feat, pcds_ind, output_size, scale_rate, plan=None):
    """Reference signature plus an optional `plan` (ops.pool_plan / ops.pool_plan_multi) so that a caller
    who knows all pooling calls of a scan up front can build their plans in one batch."""
    return VoxelMaxPoolFunction.apply(pcds_feat, pcds_ind, output_size, scale_rate, plan)
