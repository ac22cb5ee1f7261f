"""Drop-in for the pybind module `point_deep.cuda_kernel` (deep_point/src/point_deep_cuda.cpp:22-62),
so the reference's own deep_point/__init__.py runs unmodified on top of the sm_100a kernels.

The reference passes grid sizes/strides/scales as DEVICE tensors; sizes are taken from
`voxel_out.shape`, and the two scale factors are read by the plan kernel from the device tensor
itself (smos_pool_plan_desc.scale_dev): no device->host copy, no synchronisation, so the
unmodified reference path can be captured into a CUDA graph."""
import torch

from .. import ops


def _check_input(t, name):
    # CHECK_CUDA / CHECK_CONTIGUOUS (point_deep_cuda.cpp:11-13)
    if not t.is_cuda:
        raise RuntimeError("%s must be a CUDA tensor" % name)
    if not t.is_contiguous():
        raise RuntimeError("%s must be contiguous" % name)


def _plan(pcds_ind, voxel_out, voxel_max_idx, scale_rate):
    out_size = tuple(voxel_out.shape[2:])
    return ops.pool_plan(pcds_ind, out_size, None, idx_out=voxel_max_idx, idx_batch_stride=voxel_out.stride(0),
                         scale_dev=scale_rate.detach().to(torch.float32))


def voxel_maxpooling_forward(pcds_feat, pcds_ind, voxel_out, voxel_max_idx, voxel_out_size, voxel_out_stride,
                             output_size, scale_rate):
    for n, t in (("pcds_feat", pcds_feat), ("pcds_ind", pcds_ind), ("voxel_out", voxel_out),
                 ("voxel_max_idx", voxel_max_idx), ("voxel_out_size", voxel_out_size),
                 ("voxel_out_stride", voxel_out_stride), ("output_size", output_size), ("scale_rate", scale_rate)):
        _check_input(t, n)
    plan = _plan(pcds_ind, voxel_out, voxel_max_idx, scale_rate)
    ops.voxel_maxpool_forward(pcds_feat, plan, out=voxel_out)


def voxel_maxpooling_backward(pcds_feat, pcds_ind, voxel_out, voxel_max_idx, grad_pcds_feat, grad_voxel_out,
                              voxel_out_size, voxel_out_stride, output_size, scale_rate):
    for n, t in (("pcds_feat", pcds_feat), ("pcds_ind", pcds_ind), ("voxel_out", voxel_out),
                 ("voxel_max_idx", voxel_max_idx), ("grad_pcds_feat", grad_pcds_feat),
                 ("grad_voxel_out", grad_voxel_out), ("voxel_out_size", voxel_out_size),
                 ("voxel_out_stride", voxel_out_stride), ("output_size", output_size), ("scale_rate", scale_rate)):
        _check_input(t, n)
    plan = _plan(pcds_ind, voxel_out, None, scale_rate)
    ops.voxel_maxpool_backward(pcds_feat, plan, voxel_out, grad_voxel_out, grad_feat=grad_pcds_feat)
