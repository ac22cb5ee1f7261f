"""Synthetic SemanticKITTI-shaped scans for tests and benchmarks (no dataset is available).

A 64-beam spinning LiDAR over a ground plane with random obstacles, stored ring-major / azimuth-minor
like the KITTI .bin files the reference streams in file order (datasets/data_StreamMOS.py:520-574 keeps
that order, range-filters, then pads with x=y=-1000, z=-4000 up to frame_point_num). Coordinates are
quantised exactly as the reference loader does (datasets/utils.py:151-192).
"""
import numpy as np

BEV_SHAPE = (512, 512, 30)            # config/StreamMOS.py:18
RV_SHAPE = (64, 2048)                 # :19
RANGE_X, RANGE_Y, RANGE_Z = (-50.0, 50.0), (-50.0, 50.0), (-4.0, 2.0)  # :13-15
RV_THETA = (-25.0, 3.0)               # :12
PAD_XY, PAD_Z = -1000.0, -4000.0      # data_StreamMOS.py:570-571


def lidar_scan(rng, n_points=120000, n_beams=64, ego_shift=(0.0, 0.0)):
    """-> (n_points, 4) float32 xyz+intensity in scan order, all inside the crop box, then pads."""
    n_az = n_points // n_beams
    elev = np.deg2rad(np.linspace(2.0, -24.8, n_beams))
    az = np.linspace(-np.pi, np.pi, n_az, endpoint=False)
    # obstacle range per azimuth: piecewise-constant "buildings/cars" between 4 m and 48 m
    n_seg = 48
    seg_r = rng.uniform(4.0, 48.0, n_seg)
    seg_r[rng.uniform(0, 1, n_seg) < 0.35] = 48.0
    obst = np.repeat(seg_r, int(np.ceil(n_az / n_seg)))[:n_az]
    sensor_h = 1.73
    pts = np.empty((n_beams, n_az, 4), np.float32)
    for b in range(n_beams):
        if elev[b] < -1e-3:
            r_ground = sensor_h / np.tan(-elev[b])
        else:
            r_ground = 1e9
        r = np.minimum(r_ground, obst) * (1.0 + rng.normal(0, 0.004, n_az))
        r = np.clip(r, 1.5, 49.0 / max(np.cos(elev[b]), 1e-3) * 0.999)
        a = az + rng.normal(0, 2e-4, n_az)
        pts[b, :, 0] = r * np.cos(elev[b]) * np.cos(a) + ego_shift[0]
        pts[b, :, 1] = r * np.cos(elev[b]) * np.sin(a) + ego_shift[1]
        pts[b, :, 2] = np.clip(r * np.sin(elev[b]), -3.9, 1.9)
        pts[b, :, 3] = rng.uniform(0, 1, n_az)
    pts = pts.reshape(-1, 4)
    # range filter of the loader (datasets/utils.py:107-113), then pad to n_points
    keep = ((pts[:, 0] >= RANGE_X[0]) & (pts[:, 0] < RANGE_X[1]) & (pts[:, 1] >= RANGE_Y[0]) &
            (pts[:, 1] < RANGE_Y[1]) & (pts[:, 2] >= RANGE_Z[0]) & (pts[:, 2] < RANGE_Z[1]))
    # drop ~1% of returns (no-return beams) so that the tensor carries real padding
    keep &= rng.uniform(0, 1, len(pts)) > 0.01
    valid = pts[keep]
    out = np.full((n_points, 4), PAD_XY, np.float32)
    out[:, 2] = PAD_Z
    out[: len(valid)] = valid
    return out, len(valid)


def lidar_raw_scan(rng, n_beams=64, n_az=1860):
    """-> (n, 4) float32 RAW scan in its own sensor frame, as read from a KITTI .bin (no range filter, no padding):
    ring-major / azimuth-minor, obstacles out to 58 m (so the loader's range filter has something to remove), ~1 % of
    the beams without a return."""
    elev = np.deg2rad(np.linspace(2.0, -24.8, n_beams))
    az = np.linspace(-np.pi, np.pi, n_az, endpoint=False)
    n_seg = 48
    seg_r = rng.uniform(4.0, 58.0, n_seg)
    seg_r[rng.uniform(0, 1, n_seg) < 0.3] = 58.0
    obst = np.repeat(seg_r, int(np.ceil(n_az / n_seg)))[:n_az]
    sensor_h = 1.73
    pts = np.empty((n_beams, n_az, 4), np.float32)
    for b in range(n_beams):
        r_ground = sensor_h / np.tan(-elev[b]) if elev[b] < -1e-3 else 1e9
        r = np.minimum(r_ground, obst) * (1.0 + rng.normal(0, 0.004, n_az))
        r = np.maximum(r, 1.5)
        a = az + rng.normal(0, 2e-4, n_az)
        pts[b, :, 0] = r * np.cos(elev[b]) * np.cos(a)
        pts[b, :, 1] = r * np.cos(elev[b]) * np.sin(a)
        pts[b, :, 2] = r * np.sin(elev[b])
        pts[b, :, 3] = rng.uniform(0, 1, n_az)
    pts = pts.reshape(-1, 4)
    return np.ascontiguousarray(pts[rng.uniform(0, 1, len(pts)) > 0.01])


def stream_pose_diffs(t_frames=3, step=(0.62, 0.03, 0.0), yaw=0.004):
    """pose_diff of the frames of a window as the loader computes them (datasets/data_StreamMOS.py:427-447):
    inv(pose_cur).dot(pose_{cur - ht}) for a vehicle that advances by the same rigid motion every scan — the same
    T matrices for every scan of the stream (ht = 0 is only NEARLY the identity, as in the loader)."""
    c, s_ = np.cos(yaw), np.sin(yaw)
    delta = np.eye(4)
    delta[:2, :2] = [[c, -s_], [s_, c]]
    delta[:3, 3] = step
    poses = [np.eye(4)]
    for _ in range(16):
        poses.append(poses[-1].dot(delta))
    cur = 12
    cur_inv = np.linalg.inv(poses[cur])
    return [cur_inv.dot(poses[cur - ht]) for ht in range(t_frames)]


def align_filter_pad(raw, pose_diff, n_out):
    """The loader's per-frame steps (datasets/data_StreamMOS.py:515-574 with utils.Trans / filter_pcds_mask of
    datasets/utils.py:107-126) in numpy, as the host does them today: -> (n_out, 4) float32, number of valid points."""
    pc = raw.copy()
    if pose_diff is not None:
        tmp = pc[:, :4].T.copy()
        tmp[-1] = 1
        tmp = np.asarray(pose_diff, np.float64).dot(tmp).T
        pc[:, :3] = tmp[:, :3]
    keep = ((pc[:, 0] >= RANGE_X[0]) & (pc[:, 0] < RANGE_X[1]) & (pc[:, 1] >= RANGE_Y[0]) & (pc[:, 1] < RANGE_Y[1]) &
            (pc[:, 2] >= RANGE_Z[0]) & (pc[:, 2] < RANGE_Z[1]))
    pc = pc[keep]
    assert len(pc) < n_out, "frame does not fit (the loader asserts pad_length > 0)"
    out = np.full((n_out, 4), PAD_XY, np.float32)
    out[:, 2] = PAD_Z
    out[: len(pc)] = pc
    return out, len(pc)


def quantize_bev(pcds):
    """datasets/utils.py:151-169 Quantize with the config ranges -> (N, 3) float32 (x_quan, y_quan, z_quan)."""
    d = [np.float32((r[1] - r[0]) / s) for r, s in zip((RANGE_X, RANGE_Y, RANGE_Z), BEV_SHAPE)]
    q = np.stack([(pcds[:, 0] - np.float32(RANGE_X[0])) / d[0], (pcds[:, 1] - np.float32(RANGE_Y[0])) / d[1],
                  (pcds[:, 2] - np.float32(RANGE_Z[0])) / d[2]], -1)
    return q.astype(np.float32)


def quantize_sphere(pcds):
    """datasets/utils.py:172-192 SphereQuantize -> (N, 2) float32 (theta_quan, phi_quan)."""
    H, W = RV_SHAPE
    phi_r = (-np.pi, np.pi)
    th_r = (RV_THETA[0] * np.pi / 180.0, RV_THETA[1] * np.pi / 180.0)
    dphi = (phi_r[1] - phi_r[0]) / W
    dth = (th_r[1] - th_r[0]) / H
    x, y, z = pcds[:, 0].astype(np.float64), pcds[:, 1].astype(np.float64), pcds[:, 2].astype(np.float64)
    d = np.sqrt(x * x + y * y + z * z) + 1e-12
    phi = phi_r[1] - np.arctan2(x, y)
    theta = th_r[1] - np.arcsin(z / d)
    return np.stack((theta / dth, phi / dphi), -1).astype(np.float32)


def make_scan(seed, n_points=120000, t_frames=3):
    """One model input: T pose-aligned frames. Returns dict of float32 arrays shaped like the reference batch
    (models/StreamMOS.py:86-93): xyzi (T, N, 4), pcds_coord (T, N, 3, 1), pcds_sphere_coord (T, N, 2, 1)."""
    rng = np.random.default_rng(seed)
    frames, nvalid = [], []
    for t in range(t_frames):
        p, nv = lidar_scan(rng, n_points, ego_shift=(0.6 * t, 0.05 * t))
        frames.append(p)
        nvalid.append(nv)
    xyzi = np.stack(frames)
    coord = np.stack([quantize_bev(f) for f in frames])[..., None]
    sphere = np.stack([quantize_sphere(f) for f in frames])[..., None]
    return dict(xyzi=xyzi, pcds_coord=coord, pcds_sphere_coord=sphere, n_valid=np.array(nvalid))


def synthetic_boxes(rng, k=32):
    """K object AABBs (float32 corners' lo/hi) inside the crop box, car-sized."""
    c = np.stack([rng.uniform(-40, 40, k), rng.uniform(-40, 40, k), rng.uniform(-1.6, -0.6, k)], -1)
    half = np.stack([rng.uniform(0.8, 2.5, k), rng.uniform(0.8, 2.5, k), rng.uniform(0.5, 1.0, k)], -1)
    return (c - half).astype(np.float32), (c + half).astype(np.float32)
