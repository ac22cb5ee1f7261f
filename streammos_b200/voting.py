"""Long-term-memory voting with the reference's function signatures (voxel_voting.py:38-91,
voxel_instance_voting.py:78-135,169-187). The reference defines these inside scripts that run
argparse at import time, so they were never importable; here they are a module.

All functions take CUDA tensors and launch the sm_100a kernels; int64 in / int64 out as in the
reference."""
import torch

from . import ops


def _dims(size, scale):
    scale_xy, scale_z = scale[0], scale[1]
    return size[0] // scale_xy, size[1] // scale_xy, size[2] // scale_z


def Quantize(pcds, range_x=(-40, 62.4), range_y=(-40, 40), range_z=(-3, 5), size=(512, 512, 20)):
    """(P, >=3) float32 -> (P, 3) float32 quantised coordinates, (x - min) / d in fp32 with IEEE
    division (voxel_voting.py:77-91). The caller casts with .to(torch.int64) (truncation)."""
    dx = (range_x[1] - range_x[0]) / size[0]
    dy = (range_y[1] - range_y[0]) / size[1]
    dz = (range_z[1] - range_z[0]) / size[2]
    return ops.quantize(pcds, (range_x[0], range_y[0], range_z[0]), (dx, dy, dz))


def determine_voxel_labels(voxel_coords, semantic_labels, size, scale=[1, 1], num_classes=None):
    """(P, 3) int64 coords + (P,) int64 labels -> (x_max, y_max, z_max) int64 majority label per voxel;
    ties -> lowest class, empty -> 0 (voxel_voting.py:55-75).

    `num_classes=None` reproduces the reference's `labels.max().item() + 1` (one device sync);
    passing it (3 for StreamMOS) keeps the call asynchronous. The result is identical either way."""
    if num_classes is None:
        num_classes = int(semantic_labels.max().item()) + 1
    return ops.vote_voxel_labels(voxel_coords, semantic_labels, _dims(size, scale), num_classes)


def get_point_labels_from_voxel_labels(new_voxel_coords, voxel_labels, size, scale=[1, 1]):
    """Bounds-mask + gather of the voxel label of each point; out-of-range -> 0 (voxel_voting.py:38-53).
    The bounds are size // scale, the linear index uses voxel_labels' own dims, as in the reference."""
    xm, ym, zm = _dims(size, scale)
    X, Y, Z = (int(s) for s in voxel_labels.shape)
    if (xm, ym, zm) != (X, Y, Z):
        # the reference would index voxel_labels with its own strides but mask with size//scale;
        # only the consistent case occurs in the scripts
        raise NotImplementedError("voxel_labels shape must equal size // scale")
    return ops.vote_point_labels(new_voxel_coords, voxel_labels, (X, Y, Z))


def instance_vote_counts(local_map_points, local_map_prediction, cluster_corners):
    """Vote block of cluster() (voxel_instance_voting.py:169-187) for K clusters at once.

    cluster_corners: (K, 8, 3) AABB corners (after the +0.2 z-floor lift, :173-175).
    Returns (static_points_num (K,), dynamic_points_num (K,), cluster_label (K,)) int64 where
    dynamic_points_num counts 2 per dynamic point and label = 2 if dynamic > static else 1."""
    corners = cluster_corners.to(torch.float32)
    lo = corners.min(dim=1).values
    hi = corners.max(dim=1).values
    sums = ops.instance_vote(local_map_points, local_map_prediction, lo, hi)
    stat, dyn = sums[:, 0], sums[:, 1]
    label = torch.where(dyn > stat, torch.full_like(stat, 2), torch.full_like(stat, 1))
    return stat, dyn, label
