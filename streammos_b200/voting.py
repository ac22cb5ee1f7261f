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


def Quantize(pcds, range_x=(-40, 62.4), range_y=(-40, 40), range_z=(-3, 5), size=(512, 512, 20), arithmetic="ieee"):
    """(P, >=3) float32 -> (P, 3) float32 quantised coordinates, (x - min) / d in fp32 (voxel_voting.py:77-91). The
    caller casts with .to(torch.int64) (truncation).
    arithmetic="ieee" (default): IEEE division — numpy, torch on the CPU, the oracle, the golden fixtures.
    arithmetic="torch_cuda": what torch computes for the reference's expression on a CUDA device (tensor / Python scalar
    = tensor * float32(1 / scalar)), i.e. what voxel_voting.py itself produces when it runs on a GPU; the two differ in
    the last bit of some quotients (tests/test_gpu_parity.py checks this mode against torch on the device)."""
    dx = (range_x[1] - range_x[0]) / size[0]
    dy = (range_y[1] - range_y[0]) / size[1]
    dz = (range_z[1] - range_z[0]) / size[2]
    return ops.quantize(pcds, (range_x[0], range_y[0], range_z[0]), (dx, dy, dz), arithmetic)


def quantize_staged(ring_points, ring_pred, range_x, range_y, range_z, size, new_points=None, new_pred=None,
                    cur_slot=0, hist_slot=-1, want_q=True, crop_eps=None):
    """The three tensors voxel_voting.py:234-241 builds with `Quantize(...)`, `.to(torch.int64)` and
    `local_map_prediction.to(torch.int64)`, from ONE kernel: (q float32 (P, 3), voxel_coords int64 (P, 3),
    semantic_labels int64 (P,)) for the P = S*N points of a resident long-term memory ring (ring_points (S, N, >=3)
    float32, ring_pred (S, N) uint8). Same values as the three separate calls. With new_points / new_pred the new scan is
    inserted into the ring first (slot cur_slot -> hist_slot, new scan -> cur_slot). `want_q=False` skips the float32
    tensor (q is None): in the script it is a temporary that dies at the cast (:240), 12 bytes per point of writes.
    `crop_eps` (the script's 1e-4): apply transforms.Crop(fov -/+ eps) of voxel_voting.py:225-231 — a point outside
    the open box (range_min + eps, range_max - eps) keeps its slot but gets coords (-1, -1, -1), so it casts no vote and
    reads no voxel label, exactly as if the crop had removed it. None: pre-cropped input, as the bare functions assume."""
    import numpy as np
    mins = (range_x[0], range_y[0], range_z[0])
    deltas = tuple(float(np.float32((r[1] - r[0]) / s)) for r, s in zip((range_x, range_y, range_z), size))
    crop = None
    if crop_eps is not None:  # thresholds as the float32 comparison of utils/transforms.py:155-157 sees them
        crop = ([float(np.float32(r[0] + crop_eps)) for r in (range_x, range_y, range_z)],
                [float(np.float32(r[1] - crop_eps)) for r in (range_x, range_y, range_z)])
    return ops.vote_stage(ring_points, ring_pred, mins, deltas, new_points, new_pred, cur_slot, hist_slot, want_q, crop)


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


def dbscan_fit_predict(points, eps=0.3, min_samples=5):
    """sklearn.cluster.DBSCAN(eps, min_samples).fit_predict(points[:, :3]) on the device (the call at
    voxel_instance_voting.py:150-153): (M, >=3) f32 CUDA -> (M,) int32 labels, -1 = noise, same numbering."""
    m = int(points.size(0))
    st = ops.cluster_boxes(points, torch.full((m,), 2, dtype=torch.int32, device=points.device), eps, min_samples)
    return st["fg_label"][:m]


def cluster(current_points_orin, current_pred_result_orin, current_pred_bf_result_orin, local_map_points,
            local_map_prediction, eps=0.3, min_samples=5):
    """cluster() of voxel_instance_voting.py:144-193 on the device, same argument order: DBSCAN over the points the
    per-frame prediction calls moving (pred_bf == 2), one lifted axis-aligned box per cluster of more than 30
    points, votes of the local map inside each box, and the voted label written to the cluster's points.
    All tensors CUDA; current_pred_result_orin int64, modified in place and returned like the reference does.
    Eleven kernel launches and no host synchronisation: the number of clusters stays on the device."""
    pred = current_pred_result_orin
    if pred.dtype != torch.int64 or not pred.is_contiguous():
        raise RuntimeError("current_pred_result_orin must be a contiguous int64 tensor")
    if int(current_points_orin.size(0)) == 0:
        return pred
    st = ops.cluster_boxes(current_points_orin, current_pred_bf_result_orin, eps, min_samples)
    # no moving point (:146-147) or no cluster above the cut: counts[2] == 0 and both kernels do nothing
    sums = ops.instance_vote(local_map_points, local_map_prediction, st["box_lo"], st["box_hi"],
                             count=st["counts"][2:3])
    return ops.cluster_apply(st, sums, pred)


class StreamingVoter:
    """GPU-resident long-term memory for one scan stream: the loop body of voxel_voting.py:176-244 without the
    per-frame reload of 8 scans + predictions, the numpy pose transform, the crop and the .cuda() round trip.

    push(points (n, 4) f32 CUDA, pred (n,) uint8 CUDA, pose 4x4 float64) appends a scan; vote() returns the
    refined labels of the newest scan: history scans are pose-aligned into its frame (utils.Trans with
    pose_diff = inv(pose_cur) . pose_hist), history and current are cropped to fov +/- eps (transforms.Crop),
    quantised, voted per voxel, and points inside the crop take their voxel's label."""

    def __init__(self, frames_num_max=8, fov=((-50, -50, -4), (50, 50, 2)), eps=1e-4, size=(512, 512, 30),
                 num_classes=3):
        import numpy as np
        self.np = np
        self.frames_num_max = frames_num_max
        self.size = tuple(size)
        self.num_classes = num_classes
        # thresholds as the float32 tensor comparison of transforms.py:155-157 sees them
        self.crop_lo = [float(np.float32(fov[0][i] + eps)) for i in range(3)]
        self.crop_hi = [float(np.float32(fov[1][i] - eps)) for i in range(3)]
        self.mins = [float(fov[0][i]) for i in range(3)]
        self.deltas = [float(np.float32((fov[1][i] - fov[0][i]) / size[i])) for i in range(3)]
        self.ring = []  # (points, labels, pose) of the last frames_num_max + 1 scans, newest last

    def push(self, points, pred, pose):
        self.ring.append((points, pred, self.np.asarray(pose, dtype=self.np.float64)))
        if len(self.ring) > self.frames_num_max + 1:
            self.ring.pop(0)

    def vote(self, current=-1):
        """Refined labels of ring entry `current` (default: the newest scan) voted against every other resident
        scan. The steady-state branch of the reference (voxel_voting.py:177-194, id >= frames_num_max) is
        push() + vote(); its warm-up branch (:195-214: the first frames_num_max scans vote against each other)
        is frames_num_max push() calls followed by vote(current=i) for i in range(frames_num_max)."""
        np = self.np
        current = current % len(self.ring)
        cur_pts, cur_pred, cur_pose = self.ring[current]
        inv = np.linalg.inv(cur_pose)  # voxel_voting.py:179
        scans = [(p, l, inv.dot(pose) if j != current else None)    # :187-188 pose_diff
                 for j, (p, l, pose) in enumerate(self.ring)]
        return ops.vote_stream(scans, current, self.crop_lo, self.crop_hi, self.mins, self.deltas, self.size,
                               self.num_classes)
