"""Drop-in for the pybind module `point_deep.cuda_kernel` (deep_point/src/point_deep_cuda.cpp:22-62),
so the reference's own deep_point/__init__.py runs unmodified on top of the sm_100a kernels.

The reference passes grid sizes/strides/scales as DEVICE tensors; sizes are taken from
`voxel_out.shape`, the two scale factors need one small device->host read (the higher-level
`streammos_b200.deep_point` boundary avoids it)."""
import torch

from .. import ops


def _check_input(t, name):
    # CHECK_CUDA / CHECK_CONTIGUOUS (point_deep_cuda.cpp:11-13)
    if not t.is_cuda:
        raise RuntimeError("%s must be a CUDA tensor" % name)
    if not t.is_contiguous():
        raise RuntimeError("%s must be contiguous" % name)


def _plan(pcds_ind, voxel_out, voxel_max_idx, scale_rate):
    scale = [float(s) for s in scale_rate.detach().cpu().tolist()]
    out_size = tuple(voxel_out.shape[2:])
    return ops.pool_plan(pcds_ind, out_size, scale, idx_out=voxel_max_idx,
                         idx_batch_stride=voxel_out.stride(0))


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
