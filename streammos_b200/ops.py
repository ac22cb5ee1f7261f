"""Torch-tensor front end of the C-ABI: allocates outputs, passes raw pointers/strides and the
current CUDA stream. Nothing here computes — every function ends in a kernel launch inside
libstreammos_b200.so and raises if the tensors are not on a CUDA device."""
import ctypes

import torch

from . import _lib


_LAUNCHES = [0]


def reset_launch_count():
    _LAUNCHES[0] = 0


def launch_count():
    """Kernels of libstreammos_b200.so launched through this module since the last reset."""
    return _LAUNCHES[0]


def _count(n):
    _LAUNCHES[0] += n


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)


def _need_cuda(t, name):
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise RuntimeError("%s must be a CUDA tensor (streammos_b200 has no CPU path)" % name)


def _need_f32(t, name):
    if t.dtype != torch.float32:
        raise NotImplementedError("%s: only float32 is implemented on the B200 path (got %s)" % (name, t.dtype))


# ----------------------------------------------------------------------------------------------
# VoxelMaxPool
# ----------------------------------------------------------------------------------------------
class PoolPlan:
    """Per-(coordinate array, grid, scale) bucketing of points by output tile.

    Re-usable across every pooling call that shares the coordinates (forward and backward)."""

    __slots__ = ("buf", "B", "N", "H", "W", "voxel_max_idx", "scale", "taps")

    def __init__(self, buf, B, N, H, W, voxel_max_idx, scale=None, taps=None):
        self.buf, self.B, self.N, self.H, self.W, self.voxel_max_idx = buf, B, N, H, W, voxel_max_idx
        self.scale = None if scale is None else (float(scale[0]), float(scale[1]))
        self.taps = taps  # BilinearSample sampling records for this grid / scale (48 bytes per point), or None


def _plan_inputs(pcds_ind, output_size, scale_rate):
    _need_cuda(pcds_ind, "pcds_ind")
    _need_f32(pcds_ind, "pcds_ind")
    if pcds_ind.dim() == 4:
        if pcds_ind.size(3) != 1:
            raise RuntimeError("pcds_ind last dimension must be 1")
        ind = pcds_ind[..., 0]
    else:
        ind = pcds_ind
    if ind.dim() != 3 or ind.size(2) != 2 or len(output_size) != 2 or len(scale_rate) != 2:
        raise NotImplementedError("only 2-D grids (D == 2) are implemented — the shapes StreamMOS uses")
    return ind, int(ind.size(0)), int(ind.size(1)), int(output_size[0]), int(output_size[1])


def pool_plan_multi(specs, gather_taps=None):
    """Build several plans with four kernel launches in total. `specs`: list of (pcds_ind, output_size,
    scale_rate) — e.g. the five pooling calls of one scan. Returns a list of PoolPlan (views of one buffer).
    `gather_taps`: optional list of bools — plans that should also carry the BilinearSample sampling state of their
    points for their own grid and scale (used by bilinear_gather_forward(..., order=plan) when the geometry matches)."""
    lib = _lib.load()
    items, offsets, total = [], [], 0
    want = list(gather_taps) if gather_taps is not None else [False] * len(specs)
    for (pcds_ind, output_size, scale_rate), wt in zip(specs, want):
        ind, B, N, H, W = _plan_inputs(pcds_ind, output_size, scale_rate)
        nbytes = lib.smos_pool_plan_bytes(B, N, H, W)
        if nbytes < 0:
            _lib.check(int(nbytes), "smos_pool_plan_bytes")
        tbytes = int(lib.smos_gather_taps_bytes(B, N)) if wt else 0
        items.append((ind, B, N, H, W, scale_rate, int(nbytes), tbytes))
        offsets.append(total)
        total += (int(nbytes) + 255) // 256 * 256 + (tbytes + 255) // 256 * 256
    device = items[0][0].device
    big = torch.empty(total, dtype=torch.uint8, device=device)
    descs = (_lib.PoolPlanDesc * len(items))()
    plans = []
    for d, (ind, B, N, H, W, scale_rate, nbytes, tbytes), off in zip(descs, items, offsets):
        buf = big[off:off + nbytes]
        toff = off + (nbytes + 255) // 256 * 256
        taps = big[toff:toff + tbytes] if tbytes else None
        d.pcds_ind, d.B, d.N = ind.data_ptr(), B, N
        d.ind_sb, d.ind_sn, d.ind_sd = ind.stride(0), ind.stride(1), ind.stride(2)
        d.H, d.W, d.scale_h, d.scale_w = H, W, float(scale_rate[0]), float(scale_rate[1])
        d.voxel_max_idx, d.idx_batch_stride, d.plan = None, 0, buf.data_ptr()
        d.gather_taps = taps.data_ptr() if taps is not None else None
        d.scale_dev = None
        plans.append(PoolPlan(buf, B, N, H, W, None, scale_rate, taps))
    with torch.cuda.device(device):
        rc = lib.smos_pool_plan_build_multi(descs, len(items), _stream())
    _lib.check(rc, "smos_pool_plan_build_multi")
    _count(4)  # zero counts + cell index + cell allocation + scatter
    return plans


def pool_plan(pcds_ind, output_size, scale_rate, idx_out=None, idx_batch_stride=0, scale_dev=None):
    """pcds_ind (B, N, 2[, 1]) float32 -> PoolPlan. `idx_out` (B, N) int64 receives the reference's
    voxel_max_idx side product if given (deep_point/__init__.py:27). `scale_dev`: a (2,) float32 CUDA tensor that
    overrides `scale_rate` and is read by the kernels on the device (the reference hands the scales to
    point_deep.cuda_kernel as a device tensor; no device->host copy, no synchronisation)."""
    if scale_dev is not None:
        _need_cuda(scale_dev, "scale_rate")
        _need_f32(scale_dev, "scale_rate")
        if scale_dev.numel() != 2 or not scale_dev.is_contiguous():
            raise NotImplementedError("only 2-D grids (D == 2) are implemented — the shapes StreamMOS uses")
        scale_rate = (0.0, 0.0)
    ind, B, N, H, W = _plan_inputs(pcds_ind, output_size, scale_rate)
    lib = _lib.load()
    nbytes = lib.smos_pool_plan_bytes(B, N, H, W)
    if nbytes < 0:
        _lib.check(int(nbytes), "smos_pool_plan_bytes")
    buf = torch.empty(int(nbytes), dtype=torch.uint8, device=ind.device)
    if idx_out is not None:
        _need_cuda(idx_out, "voxel_max_idx")
        assert idx_out.dtype == torch.int64 and idx_out.is_contiguous() and idx_out.numel() == B * N
    d = (_lib.PoolPlanDesc * 1)()
    d[0].pcds_ind, d[0].B, d[0].N = ind.data_ptr(), B, N
    d[0].ind_sb, d[0].ind_sn, d[0].ind_sd = ind.stride(0), ind.stride(1), ind.stride(2)
    d[0].H, d[0].W, d[0].scale_h, d[0].scale_w = H, W, float(scale_rate[0]), float(scale_rate[1])
    d[0].voxel_max_idx = idx_out.data_ptr() if idx_out is not None else None
    d[0].idx_batch_stride, d[0].plan, d[0].gather_taps = int(idx_batch_stride), buf.data_ptr(), None
    d[0].scale_dev = scale_dev.data_ptr() if scale_dev is not None else None
    with torch.cuda.device(ind.device):
        rc = lib.smos_pool_plan_build_multi(d, 1, _stream())
    _lib.check(rc, "smos_pool_plan_build")
    _count(4)  # zero counts + cell index + cell allocation + scatter
    return PoolPlan(buf, B, N, H, W, idx_out, None if scale_dev is not None else scale_rate)


def cached_pool_plan(pcds_ind, output_size, scale_rate):
    """pool_plan() through the plan cache (plan_cache.py): calls that pass the same coordinate tensor, grid and scale
    — the reference's pools and gathers of one scan do — share one plan."""
    from . import plan_cache
    ind, B, N, H, W = _plan_inputs(pcds_ind, output_size, scale_rate)
    return plan_cache.get(ind, (H, W), scale_rate, lambda: pool_plan(ind, (H, W), scale_rate),
                          lambda geos: pool_plan_multi([(ind, (g[0], g[1]), (g[2], g[3])) for g in geos]))


def _feat3(pcds_feat):
    if pcds_feat.dim() == 4:
        if pcds_feat.size(3) != 1:
            raise RuntimeError("pcds_feat last dimension must be 1")
        return pcds_feat[..., 0]
    return pcds_feat


POOL_STAGE_REDUCE, POOL_STAGE_COMBINE, POOL_STAGE_WRITE, POOL_STAGE_ALL = 1, 2, 4, 7


def voxel_maxpool_forward(pcds_feat, plan, out=None, stages=POOL_STAGE_ALL, workspace=None):
    """pcds_feat (B, C, N[, 1]) float32, any strides -> (B, C, H, W) NCHW-contiguous.
    `stages` / `workspace` let a benchmark re-launch a single stage of a call that already ran once."""
    _need_cuda(pcds_feat, "pcds_feat")
    _need_f32(pcds_feat, "pcds_feat")
    f = _feat3(pcds_feat)
    B, C, N = (int(s) for s in f.shape)
    if B != plan.B or N != plan.N:
        raise RuntimeError("pcds_feat (B=%d, N=%d) does not match the plan (B=%d, N=%d)" % (B, N, plan.B, plan.N))
    if out is None:
        out = torch.empty((B, C, plan.H, plan.W), dtype=torch.float32, device=f.device)
    else:
        assert out.is_contiguous() and out.shape == (B, C, plan.H, plan.W) and out.dtype == torch.float32
    lib = _lib.load()
    ws = workspace
    if ws is None:
        ws = torch.empty(int(lib.smos_pool_workspace_bytes(B, C, N)), dtype=torch.uint8, device=f.device)
    point_major = f.stride(1) == 1 and C > 1
    with torch.cuda.device(f.device):
        rc = lib.smos_voxel_maxpool_forward_stages(_ptr(f), B, C, N, f.stride(0), f.stride(1), f.stride(2), plan.H,
                                                   plan.W, _ptr(plan.buf), _ptr(ws), _ptr(out), int(stages), _stream())
    _lib.check(rc, "smos_voxel_maxpool_forward")
    import os
    mode = os.environ.get("SMOS_POOL_FOLD", "2")  # default: folded inside the writer launch unless the output exceeds L2
    big_out = B * C * plan.H * plan.W * 4 > (96 << 20)
    separate_combine = (mode == "0" or (mode == "2" and big_out)) and B * N >= 32
    _count((1 if point_major else 2) * bool(stages & 1) + (bool(stages & 2) and separate_combine) + bool(stages & 4))
    return out


def pool_workspace(B, C, N, device):
    return torch.empty(int(_lib.load().smos_pool_workspace_bytes(B, C, N)), dtype=torch.uint8, device=device)


def voxel_maxpool_backward(pcds_feat, plan, voxel_out, grad_voxel_out, grad_feat=None):
    _need_cuda(pcds_feat, "pcds_feat")
    _need_f32(pcds_feat, "pcds_feat")
    f = _feat3(pcds_feat)
    B, C, N = (int(s) for s in f.shape)
    grad_voxel_out = grad_voxel_out.contiguous()
    assert voxel_out.is_contiguous()
    if grad_feat is None:
        grad_feat = torch.empty(pcds_feat.shape, dtype=torch.float32, device=f.device)
    g = _feat3(grad_feat)
    with torch.cuda.device(f.device):
        rc = _lib.load().smos_voxel_maxpool_backward(_ptr(f), B, C, N, f.stride(0), f.stride(1), f.stride(2),
                                                     plan.H, plan.W, _ptr(plan.buf), _ptr(voxel_out),
                                                     _ptr(grad_voxel_out), _ptr(g), g.stride(0), g.stride(1),
                                                     g.stride(2), _stream())
    _lib.check(rc, "smos_voxel_maxpool_backward")
    _count(1)
    return grad_feat


# ----------------------------------------------------------------------------------------------
# BilinearSample
# ----------------------------------------------------------------------------------------------
def bilinear_gather_forward(grid_feat, grid_coord, scale_rate, point_major_out=False, order=None):
    """grid_feat (B, C, H, W) float32 (NCHW or channels_last), grid_coord (B, N, 2, S) float32
    -> (B, C, N, S). With point_major_out the result is channels_last-strided (each point's C
    features contiguous), the layout the pooling kernel reads fastest.
    `order`: a PoolPlan built from the same `grid_coord` (any grid/scale): points are visited in its cell
    order (same values, far fewer cache lines per warp load for BEV coordinates); S == 1, point-major only."""
    _need_cuda(grid_feat, "grid_feat")
    _need_cuda(grid_coord, "grid_coord")
    _need_f32(grid_feat, "grid_feat")
    _need_f32(grid_coord, "grid_coord")
    B, C, H, W = (int(s) for s in grid_feat.shape)
    if grid_coord.dim() != 4 or grid_coord.size(0) != B or grid_coord.size(2) != 2:
        raise RuntimeError("grid_coord must be (B, N, 2, S)")
    N, S = int(grid_coord.size(1)), int(grid_coord.size(3))
    if S == 1:
        co = grid_coord[..., 0]  # (B, N, 2)
    else:
        co = grid_coord.permute(0, 1, 3, 2).reshape(B, N * S, 2)
    NP = N * S
    if point_major_out:
        out = torch.empty((B, N, S, C), dtype=torch.float32, device=grid_feat.device).permute(0, 3, 1, 2)
    else:
        out = torch.empty((B, C, N, S), dtype=torch.float32, device=grid_feat.device)
    o_sb, o_sc = out.stride(0), out.stride(1)
    o_sn = out.stride(3)  # flattened (n, s) index advances by the stride of s
    if order is not None and not (S == 1 and point_major_out):
        order = None
    if order is not None and (order.B, order.N) != (B, N):
        raise RuntimeError("order plan (B=%d, N=%d) does not match grid_coord (B=%d, N=%d)" % (order.B, order.N, B, N))
    use_taps = (order is not None and order.taps is not None and (order.H, order.W) == (H, W) and
                order.scale == (float(scale_rate[0]), float(scale_rate[1])))
    with torch.cuda.device(grid_feat.device):
        args = (_ptr(grid_feat), B, C, H, W, grid_feat.stride(0), grid_feat.stride(1), grid_feat.stride(2),
                grid_feat.stride(3), _ptr(co), NP, co.stride(0), co.stride(1), co.stride(2), float(scale_rate[0]),
                float(scale_rate[1]), _ptr(out), o_sb, o_sc, o_sn)
        if use_taps:  # the plan carries the sampling state of exactly this gather: no coordinates needed
            rc = _lib.load().smos_bilinear_gather_forward_taps(*args[:9], _ptr(order.taps), NP, _ptr(out), o_sb, o_sc,
                                                               o_sn, _stream())
        elif order is None:
            rc = _lib.load().smos_bilinear_gather_forward(*args, _stream())
        else:
            rc = _lib.load().smos_bilinear_gather_forward_ordered(*args, _ptr(order.buf), order.H, order.W, _stream())
    _lib.check(rc, "smos_bilinear_gather_forward")
    _count(1)
    return out


def bilinear_gather_backward(grad_out, grid_coord, scale_rate, H, W):
    """grad_out (B, C, N, S) -> grad wrt grid_feat (B, C, H, W)."""
    _need_cuda(grad_out, "grad_out")
    _need_f32(grad_out, "grad_out")
    B, C, N, S = (int(s) for s in grad_out.shape)
    if S == 1:
        co = grid_coord[..., 0]
        go = grad_out[..., 0]
    else:
        co = grid_coord.permute(0, 1, 3, 2).reshape(B, N * S, 2)
        go = grad_out.reshape(B, C, N * S)
    grad_grid = torch.empty((B, C, H, W), dtype=torch.float32, device=grad_out.device)   # every element is written
    with torch.cuda.device(grad_out.device):
        rc = _lib.load().smos_bilinear_gather_backward(
            _ptr(go), B, C, N * S, go.stride(0), go.stride(1), go.stride(2), _ptr(co), co.stride(0), co.stride(1),
            co.stride(2), float(scale_rate[0]), float(scale_rate[1]), int(H), int(W), _ptr(grad_grid), _stream())
    _lib.check(rc, "smos_bilinear_gather_backward")
    _count(1)
    return grad_grid


# ----------------------------------------------------------------------------------------------
# Model input tensors from raw scans (next: SURVEY 8f rank 2, the exact part)
# ----------------------------------------------------------------------------------------------
def form_batch(points, range_x, range_y, range_z, size, x_sign=1.0, y_sign=1.0):
    """points (T, N, >=4) float32 CUDA raw x, y, z, intensity (range filtered and padded as the loader does) ->
    (pcds_xyzi (T, 7, N, 1), pcds_coord (T, N, 3, 1)): Quantize + make_point_feat of the loader's form_batch,
    bit-exact with its numpy float32 arithmetic."""
    import numpy as np
    _need_cuda(points, "points")
    _need_f32(points, "points")
    assert points.dim() == 3 and points.size(2) >= 4 and points.stride(2) == 1 and points.stride(0) == points.size(1) * points.stride(1)
    T, N = int(points.size(0)), int(points.size(1))
    d = [float(np.float32((r[1] - r[0]) / s)) for r, s in zip((range_x, range_y, range_z), size)]
    feat = torch.empty((T, 7, N, 1), dtype=torch.float32, device=points.device)
    coord = torch.empty((T, N, 3, 1), dtype=torch.float32, device=points.device)
    with torch.cuda.device(points.device):
        rc = _lib.load().smos_form_batch(_ptr(points), T, N, points.stride(1) if N > 1 else points.size(2), float(x_sign),
                                         float(y_sign), float(range_x[0]), float(range_y[0]), float(range_z[0]), d[0],
                                         d[1], d[2], _ptr(feat), _ptr(coord), _stream())
    _lib.check(rc, "smos_form_batch")
    _count(1)
    return feat, coord


def sphere_constants(phi_range=(-180.0, 180.0), theta_range=(-25.0, 3.0), size=(64, 2048)):
    """The four float32 constants of utils.SphereQuantize (datasets/utils.py:173-180) as numpy forms them: float64
    products of the degree bounds, used as weak scalars against float32 arrays -> (phi_hi, theta_hi, dphi, dtheta)."""
    import numpy as np
    H, W = size
    phi_r = (phi_range[0] * np.pi / 180.0, phi_range[1] * np.pi / 180.0)
    th_r = (theta_range[0] * np.pi / 180.0, theta_range[1] * np.pi / 180.0)
    dphi = (phi_r[1] - phi_r[0]) / W
    dth = (th_r[1] - th_r[0]) / H
    return tuple(float(np.float32(v)) for v in (phi_r[1], th_r[1], dphi, dth))


def sphere_quantize(points, phi_range=(-180.0, 180.0), theta_range=(-25.0, 3.0), size=(64, 2048), x_sign=1.0,
                    y_sign=1.0, out=None):
    """utils.SphereQuantize (datasets/utils.py:172-192) on the device: points (T, N, >=3) float32 CUDA ->
    pcds_sphere_coord (T, N, 2, 1) = (theta_quan, phi_quan). Floating point (smos_sphere_quantize): the angles are the
    float64 arctan2 / arcsin rounded to float32, everything else is the reference's float32 sequence; within 1 ulp of
    the angle of numpy's result (include/streammos_b200.h states the bound in cells)."""
    _need_cuda(points, "points")
    _need_f32(points, "points")
    assert points.dim() == 3 and points.size(2) >= 3 and points.stride(2) == 1
    T, N = int(points.size(0)), int(points.size(1))
    assert T == 1 or points.stride(0) == N * points.stride(1), "frames must follow each other at one row stride"
    c = sphere_constants(phi_range, theta_range, size)
    if out is None:
        out = torch.empty((T, N, 2, 1), dtype=torch.float32, device=points.device)
    assert out.is_contiguous() and out.numel() == T * N * 2 and out.dtype == torch.float32
    with torch.cuda.device(points.device):
        rc = _lib.load().smos_sphere_quantize(_ptr(points), T * N, points.stride(1) if N > 1 else points.size(2),
                                              float(x_sign), float(y_sign), c[0], c[1], c[2], c[3], _ptr(out), _stream())
    _lib.check(rc, "smos_sphere_quantize")
    _count(1)
    return out


def ingest_frames(frames, range_x, range_y, range_z, n_out, pad_xy=-1000.0, pad_z=-4000.0, want_src=False,
                  out=None, workspace=None):
    """The loader steps in front of form_batch on the device (datasets/data_StreamMOS.py:515-574): per frame pose
    alignment (utils.Trans), range filter (utils.filter_pcds_mask), order-preserving compaction and padding to `n_out`
    rows. `frames`: list of (points (n_cap, >=4) f32 CUDA raw scan, n, pose) with n an int or a (1,) int32 CUDA tensor
    (rows that hold points) and pose a 4x4 / 3x4 float64 array, a (12,) float64 CUDA tensor, or None.
    -> (points (T, n_out, 4) f32, count (T,) int32 CUDA[, src (T, n_out) int32]); bit-exact with the loader."""
    import numpy as np
    lib = _lib.load()
    T = len(frames)
    dev = frames[0][0].device
    descs = (_lib.IngestFrame * T)()
    keep = []
    n_cap_max, rs = 0, None
    for d, (pts, n, pose) in zip(descs, frames):
        _need_cuda(pts, "points")
        _need_f32(pts, "points")
        if pts.dim() != 2 or pts.size(1) < 4 or pts.stride(1) != 1:
            raise RuntimeError("ingest_frames: points must be (n_cap, >=4) float32 rows")
        r = pts.stride(0) if pts.size(0) > 1 else pts.size(1)
        rs = r if rs is None else rs
        if r != rs:
            raise RuntimeError("ingest_frames: all frames must share a row stride")
        if not isinstance(n, torch.Tensor):
            n = torch.tensor([int(n)], dtype=torch.int32, device=dev)
        if n.dtype != torch.int32 or not n.is_cuda:
            raise RuntimeError("ingest_frames: n must be an int or an int32 CUDA tensor")
        if pose is not None and not isinstance(pose, torch.Tensor):
            m = np.asarray(pose, dtype=np.float64).reshape(-1, 4)[:3]
            pose = torch.from_numpy(np.ascontiguousarray(m).reshape(12)).to(dev)
        if pose is not None and (pose.dtype != torch.float64 or pose.numel() < 12 or not pose.is_cuda or
                                 not pose.is_contiguous()):
            raise RuntimeError("ingest_frames: pose must hold 12 contiguous float64 values on the device")
        keep.append((n, pose))
        d.points, d.n_dev, d.n_cap = pts.data_ptr(), n.data_ptr(), int(pts.size(0))
        d.pose_dev = pose.data_ptr() if pose is not None else None
        n_cap_max = max(n_cap_max, int(pts.size(0)))
    n_out = int(n_out)
    if out is None:
        out = torch.empty((T, n_out, 4), dtype=torch.float32, device=dev)
    count = torch.empty((T,), dtype=torch.int32, device=dev)
    src = torch.empty((T, n_out), dtype=torch.int32, device=dev) if want_src else None
    if workspace is None:
        workspace = torch.empty(int(lib.smos_ingest_workspace_bytes(T, n_cap_max, n_out)), dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        rc = lib.smos_ingest_frames(descs, T, int(rs), float(range_x[0]), float(range_x[1]), float(range_y[0]),
                                    float(range_y[1]), float(range_z[0]), float(range_z[1]), n_out, float(pad_xy),
                                    float(pad_z), _ptr(workspace), _ptr(out), _ptr(src), _ptr(count), _stream())
    _lib.check(rc, "smos_ingest_frames")
    _count(2)
    return (out, count, src) if want_src else (out, count)


# ----------------------------------------------------------------------------------------------
# PointNet stem (next: SURVEY 8f rank 4)
# ----------------------------------------------------------------------------------------------
def _stem_out(B, C2, N, device, point_major):
    import os
    if point_major and os.environ.get("SMOS_STEM_UMMA", "1") != "0" and os.environ.get("SMOS_STEM_TC", "0") == "0":
        # same shape, channels_last strides: each point's C2 features contiguous (tensor-core kernel only)
        return torch.empty((B, N, 1, C2), dtype=torch.float32, device=device).permute(0, 3, 1, 2)
    return torch.empty((B, C2, N, 1), dtype=torch.float32, device=device)


def point_stem_forward(x, bn0, w1, bn1, w2, bn2, out=None, point_major_out=False):
    """Fused eval-mode PointNetStacker(Cin, 64, pre_bn=True, stack_num=2): x (B, Cin, N[, 1]) float32 ->
    (B, 64, N, 1), contiguous or (point_major_out) with channels_last strides — the layout VoxelMaxPool reads without its
    permute stage. bn* = (alpha, beta) per-channel affines of the eval BatchNorms (bn0 may be None);
    w1 (64, Cin[, 1, 1]), w2 (64, 64[, 1, 1])."""
    _need_cuda(x, "x")
    _need_f32(x, "x")
    x3 = _feat3(x)
    B, Cin, N = (int(v) for v in x3.shape)
    for name, w in (("w1", w1), ("w2", w2)):
        _need_cuda(w, name)
        _need_f32(w, name)
    w1 = w1.reshape(w1.shape[0], -1).contiguous()
    w2 = w2.reshape(w2.shape[0], -1).contiguous()
    C1, C2 = int(w1.shape[0]), int(w2.shape[0])
    if w1.shape[1] != Cin or w2.shape[1] != C1:
        raise RuntimeError("point_stem: weight shapes do not chain (Cin=%d, w1 %s, w2 %s)" % (Cin, tuple(w1.shape), tuple(w2.shape)))
    vecs = [t for t in ((bn0 or (None, None)) + tuple(bn1) + tuple(bn2)) if t is not None]
    for v in vecs:
        _need_cuda(v, "BatchNorm affine")
        _need_f32(v, "BatchNorm affine")
    vecs = [v.contiguous() for v in vecs]
    a0, b0 = (vecs[0], vecs[1]) if bn0 is not None else (None, None)
    a1, b1, a2, b2 = vecs[-4:]
    if out is None:
        out = _stem_out(B, C2, N, x.device, point_major_out)
    else:
        assert out.shape == (B, C2, N, 1) and out.dtype == torch.float32
    with torch.cuda.device(x.device):
        rc = _lib.load().smos_point_stem_forward(_ptr(x3), B, Cin, N, x3.stride(0), x3.stride(1), x3.stride(2), _ptr(a0),
                                                 _ptr(b0), _ptr(w1), _ptr(a1), _ptr(b1), _ptr(w2), _ptr(a2), _ptr(b2),
                                                 C1, C2, _ptr(out), out.stride(0), out.stride(1), out.stride(2), _stream())
    _lib.check(rc, "smos_point_stem_forward")
    _count(1)
    return out


def point_stem_forward_raw(points, range_x, range_y, range_z, size, bn0, w1, bn1, w2, bn2, x_sign=1.0, y_sign=1.0,
                           point_major_out=False, max_ctas=0):
    """form_batch + point_stem_forward in one kernel: points (T, N, >=4) float32 CUDA raw scans ->
    (features (T, 64, N, 1), pcds_coord (T, N, 3, 1)); bit-identical to the two separate calls.
    max_ctas > 0 caps the persistent CTAs of the kernel (scheduling hint for pipelined streams, same results)."""
    import numpy as np
    _need_cuda(points, "points")
    _need_f32(points, "points")
    assert points.dim() == 3 and points.size(2) >= 4 and points.stride(2) == 1 and points.stride(0) == points.size(1) * points.stride(1)
    T, N = int(points.size(0)), int(points.size(1))
    for name, w in (("w1", w1), ("w2", w2)):
        _need_cuda(w, name)
        _need_f32(w, name)
    w1 = w1.reshape(w1.shape[0], -1).contiguous()
    w2 = w2.reshape(w2.shape[0], -1).contiguous()
    if w1.shape[1] != 7 or w2.shape[1] != w1.shape[0]:
        raise RuntimeError("point_stem_forward_raw: the stem must take the 7 loader channels")
    vecs = [t for t in ((bn0 or (None, None)) + tuple(bn1) + tuple(bn2)) if t is not None]
    for v in vecs:
        _need_cuda(v, "BatchNorm affine")
        _need_f32(v, "BatchNorm affine")
    vecs = [v.contiguous() for v in vecs]
    a0, b0 = (vecs[0], vecs[1]) if bn0 is not None else (None, None)
    a1, b1, a2, b2 = vecs[-4:]
    d = [float(np.float32((r[1] - r[0]) / s_)) for r, s_ in zip((range_x, range_y, range_z), size)]
    C1, C2 = int(w1.shape[0]), int(w2.shape[0])
    out = _stem_out(T, C2, N, points.device, point_major_out)
    coord = torch.empty((T, N, 3, 1), dtype=torch.float32, device=points.device)
    with torch.cuda.device(points.device):
        rc = _lib.load().smos_point_stem_forward_raw_capped(
            _ptr(points), T, N, points.stride(1) if N > 1 else points.size(2), float(x_sign), float(y_sign),
            float(range_x[0]), float(range_y[0]), float(range_z[0]), d[0], d[1], d[2], _ptr(a0), _ptr(b0), _ptr(w1),
            _ptr(a1), _ptr(b1), _ptr(w2), _ptr(a2), _ptr(b2), C1, C2, _ptr(coord), _ptr(out), out.stride(0), out.stride(1),
            out.stride(2), int(max_ctas), _stream())
    _lib.check(rc, "smos_point_stem_forward_raw")
    _count(1)
    return out, coord


# ----------------------------------------------------------------------------------------------
# MSDeformAttn
# ----------------------------------------------------------------------------------------------
_DT = {torch.float32: 0, torch.float64: 1}


def _msda_check(value, spatial_shapes, level_start_index, sampling_loc, attn_weight, extra=()):
    named = [("value", value), ("spatial_shapes", spatial_shapes), ("level_start_index", level_start_index),
             ("sampling_loc", sampling_loc), ("attn_weight", attn_weight)] + list(extra)
    for name, t in named:
        if not t.is_contiguous():
            raise RuntimeError("%s tensor has to be contiguous" % name)  # ms_deform_attn_cuda.cu:28-32
    for name, t in named:
        if not t.is_cuda:
            raise RuntimeError("%s must be a CUDA tensor (Not implemented on the CPU)" % name)  # :34-38
    if value.dtype not in _DT:
        raise RuntimeError("ms_deform_attn: only float32/float64 are dispatched (got %s)" % value.dtype)
    if sampling_loc.dtype != value.dtype or attn_weight.dtype != value.dtype:
        raise RuntimeError("ms_deform_attn: value, sampling_loc and attn_weight must share a dtype")
    if spatial_shapes.dtype != torch.int64 or level_start_index.dtype != torch.int64:
        raise RuntimeError("ms_deform_attn: spatial_shapes / level_start_index must be int64")
    B, S, M, D = (int(s) for s in value.shape)
    L = int(spatial_shapes.size(0))
    Q, P = int(sampling_loc.size(1)), int(sampling_loc.size(4))
    return B, S, M, D, L, Q, P


def ms_deform_attn_forward(value, spatial_shapes, level_start_index, sampling_loc, attn_weight, out=None):
    """`out`: optional (B, Lq, M*D) contiguous tensor to write into (a stream keeps its short-term memory in place: the
    second layer's output lands straight in the memory buffer the next scan reads). Must not alias `value`."""
    B, S, M, D, L, Q, P = _msda_check(value, spatial_shapes, level_start_index, sampling_loc, attn_weight)
    if out is None:
        out = torch.empty((B, Q, M * D), dtype=value.dtype, device=value.device)
    else:
        if out.dtype != value.dtype or not out.is_contiguous() or out.numel() != B * Q * M * D or not out.is_cuda:
            raise RuntimeError("out must be a contiguous (B, Lq, M*D) CUDA tensor of value's dtype")
        if out.data_ptr() == value.data_ptr():
            raise RuntimeError("out must not alias value")
    with torch.cuda.device(value.device):
        rc = _lib.load().smos_ms_deform_attn_forward(_DT[value.dtype], _ptr(value), _ptr(spatial_shapes),
                                                     _ptr(level_start_index), _ptr(sampling_loc),
                                                     _ptr(attn_weight), B, S, M, D, L, Q, P, _ptr(out), _stream())
    _lib.check(rc, "smos_ms_deform_attn_forward")
    _count(1)
    return out


def ms_deform_attn_fused_forward(value, spatial_shapes, level_start_index, sampling_offsets, attn_logits,
                                 reference_points, out=None):
    """The sampling core plus what deformattn/modules/ms_deform_attn.py:96-108 computes in front of it, in one kernel:
    softmax of `attn_logits` (B, Lq, M, L*P) over the L*P samples and sampling_locations = reference_points (B, Lq, L,
    2 | 4) + normalised `sampling_offsets` (B, Lq, M, L, P, 2). -> (B, Lq, M*D). Forward only (inference)."""
    B, S, M, D, L, Q, P = _msda_check(value, spatial_shapes, level_start_index, sampling_offsets,
                                      attn_logits.view(attn_logits.shape[0], attn_logits.shape[1], attn_logits.shape[2],
                                                       spatial_shapes.size(0), -1),
                                      extra=[("reference_points", reference_points)])
    if reference_points.dtype != value.dtype:
        raise RuntimeError("ms_deform_attn: reference_points must have value's dtype")
    rd = int(reference_points.shape[-1])
    if rd not in (2, 4) or tuple(reference_points.shape[:3]) != (B, Q, L):
        raise ValueError("Last dim of reference_points must be 2 or 4, but get {} instead.".format(rd))  # module :107
    if tuple(sampling_offsets.shape) != (B, Q, M, L, P, 2) or attn_logits.numel() != B * Q * M * L * P:
        raise RuntimeError("ms_deform_attn: sampling_offsets (B, Lq, M, L, P, 2) / attn_logits (B, Lq, M, L*P)")
    if out is None:
        out = torch.empty((B, Q, M * D), dtype=value.dtype, device=value.device)
    elif out.dtype != value.dtype or not out.is_contiguous() or out.numel() != B * Q * M * D or not out.is_cuda or \
            out.data_ptr() == value.data_ptr():
        raise RuntimeError("out must be a contiguous (B, Lq, M*D) CUDA tensor of value's dtype that does not alias value")
    with torch.cuda.device(value.device):
        rc = _lib.load().smos_ms_deform_attn_fused_forward(
            _DT[value.dtype], _ptr(value), _ptr(spatial_shapes), _ptr(level_start_index), _ptr(sampling_offsets),
            _ptr(attn_logits), _ptr(reference_points), rd, B, S, M, D, L, Q, P, _ptr(out), _stream())
    _lib.check(rc, "smos_ms_deform_attn_fused_forward")
    _count(1)
    return out


def ms_deform_attn_backward(value, spatial_shapes, level_start_index, sampling_loc, attn_weight, grad_output):
    B, S, M, D, L, Q, P = _msda_check(value, spatial_shapes, level_start_index, sampling_loc, attn_weight,
                                      extra=[("grad_output", grad_output)])
    grad_value = torch.zeros_like(value)
    grad_loc = torch.empty_like(sampling_loc)
    grad_attn = torch.empty_like(attn_weight)
    with torch.cuda.device(value.device):
        rc = _lib.load().smos_ms_deform_attn_backward(
            _DT[value.dtype], _ptr(value), _ptr(spatial_shapes), _ptr(level_start_index), _ptr(sampling_loc),
            _ptr(attn_weight), _ptr(grad_output), B, S, M, D, L, Q, P, _ptr(grad_value), _ptr(grad_loc),
            _ptr(grad_attn), _stream())
    _lib.check(rc, "smos_ms_deform_attn_backward")
    _count(1)
    return grad_value, grad_loc, grad_attn


# ----------------------------------------------------------------------------------------------
# Voting
# ----------------------------------------------------------------------------------------------
def quantize(pcds, mins, deltas, arithmetic="ieee"):
    """arithmetic="ieee": float32 IEEE division (numpy, torch on the CPU); "torch_cuda": multiplication by the float32
    reciprocal, bit-identical with torch evaluating the same expression on a CUDA device (smos_quantize_rcp)."""
    _need_cuda(pcds, "pcds")
    _need_f32(pcds, "pcds")
    assert pcds.dim() == 2 and pcds.size(1) >= 3 and pcds.stride(1) == 1
    if arithmetic not in ("ieee", "torch_cuda"):
        raise ValueError("arithmetic must be 'ieee' or 'torch_cuda'")
    P = int(pcds.size(0))
    out = torch.empty((P, 3), dtype=torch.float32, device=pcds.device)
    lib = _lib.load()
    fn = lib.smos_quantize if arithmetic == "ieee" else lib.smos_quantize_rcp
    with torch.cuda.device(pcds.device):
        rc = fn(_ptr(pcds), P, pcds.stride(0) if P > 1 else pcds.size(1), float(mins[0]), float(mins[1]), float(mins[2]),
                float(deltas[0]), float(deltas[1]), float(deltas[2]), _ptr(out), _stream())
    _lib.check(rc, "smos_quantize")
    _count(1)
    return out


def vote_stage(ring_points, ring_pred, mins, deltas, new_points=None, new_pred=None, cur_slot=0, hist_slot=-1,
               want_q=True, crop=None):
    """Quantize + the two `.to(torch.int64)` casts of voxel_voting.py:234-241 in one kernel, straight from the
    long-term memory ring: ring_points (S, N, r>=3) f32, ring_pred (S, N) u8 -> (q (S*N, 3) f32 or None,
    coords (S*N, 3) int64, labels (S*N,) int64). With new_points / new_pred the ring insert (memory_push) happens in
    the same kernel first: slot cur_slot moves to hist_slot (if >= 0) and the new scan takes cur_slot.
    `crop` = (lo, hi) float thresholds of the open box the script crops the map to before quantising
    (voxel_voting.py:225-231): points outside it get coords (-1, -1, -1) — no vote, no label."""
    _need_cuda(ring_points, "ring_points")
    _need_f32(ring_points, "ring_points")
    if ring_points.dim() != 3 or not ring_points.is_contiguous() or ring_pred.dtype != torch.uint8 or \
            not ring_pred.is_contiguous() or ring_pred.shape != ring_points.shape[:2] or not ring_pred.is_cuda:
        raise RuntimeError("vote_stage: ring_points (S, N, r) f32 / ring_pred (S, N) u8, contiguous CUDA tensors")
    S, N, r = (int(v) for v in ring_points.shape)
    if new_points is not None:
        _need_cuda(new_points, "new_points")
        _need_f32(new_points, "new_points")
        if new_points.shape != (N, r) or not new_points.is_contiguous() or new_pred.dtype != torch.uint8 or \
                new_pred.shape != (N,) or not new_pred.is_contiguous() or not new_pred.is_cuda:
            raise RuntimeError("vote_stage: new_points (N, r) f32 / new_pred (N,) u8, contiguous CUDA tensors")
    dev = ring_points.device
    q = torch.empty((S * N, 3), dtype=torch.float32, device=dev) if want_q else None
    coords = torch.empty((S * N, 3), dtype=torch.int64, device=dev)
    labels = torch.empty((S * N,), dtype=torch.int64, device=dev)
    lo = (ctypes.c_float * 3)(*[float(v) for v in crop[0]]) if crop is not None else None
    hi = (ctypes.c_float * 3)(*[float(v) for v in crop[1]]) if crop is not None else None
    with torch.cuda.device(dev):
        rc = _lib.load().smos_vote_stage(_ptr(ring_points), _ptr(ring_pred), S, N, r, _ptr(new_points), _ptr(new_pred),
                                         int(cur_slot), int(hist_slot), float(mins[0]), float(mins[1]), float(mins[2]),
                                         float(deltas[0]), float(deltas[1]), float(deltas[2]), lo, hi, _ptr(q), _ptr(coords),
                                         _ptr(labels), _stream())
    _lib.check(rc, "smos_vote_stage")
    _count(1)
    return q, coords, labels


def _vote_ws(P, X, Y, Z, C, device):
    n = _lib.load().smos_vote_workspace_bytes(P, X, Y, Z, C)
    if n < 0:
        _lib.check(int(n), "smos_vote_workspace_bytes")
    return torch.empty(int(n), dtype=torch.uint8, device=device)


def vote_voxel_labels(voxel_coords, semantic_labels, dims, num_classes):
    _need_cuda(voxel_coords, "voxel_coords")
    _need_cuda(semantic_labels, "semantic_labels")
    if voxel_coords.dtype != torch.int64 or semantic_labels.dtype != torch.int64:
        raise RuntimeError("voxel_coords / semantic_labels must be int64 (the reference API dtype)")
    voxel_coords = voxel_coords.contiguous()
    semantic_labels = semantic_labels.contiguous()
    P = int(voxel_coords.size(0))
    X, Y, Z = (int(d) for d in dims)
    ws = _vote_ws(P, X, Y, Z, int(num_classes), voxel_coords.device)
    out = torch.empty((X, Y, Z), dtype=torch.int64, device=voxel_coords.device)
    with torch.cuda.device(voxel_coords.device):
        rc = _lib.load().smos_vote_voxel_labels(_ptr(voxel_coords), _ptr(semantic_labels), P, X, Y, Z,
                                                int(num_classes), _ptr(ws), _ptr(out), _stream())
    _lib.check(rc, "smos_vote_voxel_labels")
    _count(3)  # zero + vote + per-point conversion
    return out


def vote_point_labels(new_voxel_coords, voxel_labels, dims):
    _need_cuda(new_voxel_coords, "new_voxel_coords")
    _need_cuda(voxel_labels, "voxel_labels")
    if new_voxel_coords.dtype != torch.int64 or voxel_labels.dtype != torch.int64:
        raise RuntimeError("new_voxel_coords / voxel_labels must be int64 (the reference API dtype)")
    new_voxel_coords = new_voxel_coords.contiguous()
    voxel_labels = voxel_labels.contiguous()
    Pc = int(new_voxel_coords.size(0))
    X, Y, Z = (int(d) for d in dims)
    out = torch.empty((Pc,), dtype=torch.int64, device=new_voxel_coords.device)
    with torch.cuda.device(new_voxel_coords.device):
        rc = _lib.load().smos_vote_point_labels(_ptr(new_voxel_coords), Pc, _ptr(voxel_labels), X, Y, Z, _ptr(out),
                                                _stream())
    _lib.check(rc, "smos_vote_point_labels")
    _count(1)
    return out


def vote_fused(points, labels_u8, num_current, mins, deltas, dims, num_classes=3):
    """Streaming-path voting straight from float xyz + uint8 labels (no int64 staging).
    Returns (voxel_labels uint8 (X,Y,Z), point_labels int64 (num_current,))."""
    _need_cuda(points, "points")
    _need_cuda(labels_u8, "labels")
    _need_f32(points, "points")
    assert labels_u8.dtype == torch.uint8 and labels_u8.is_contiguous()
    assert points.dim() == 2 and points.size(1) >= 3 and points.stride(1) == 1
    P = int(points.size(0))
    X, Y, Z = (int(d) for d in dims)
    ws = _vote_ws(P, X, Y, Z, int(num_classes), points.device)
    vl = torch.empty((X, Y, Z), dtype=torch.uint8, device=points.device)
    pl = torch.empty((int(num_current),), dtype=torch.int64, device=points.device)
    with torch.cuda.device(points.device):
        rc = _lib.load().smos_vote_fused(_ptr(points), P, points.stride(0) if P > 1 else points.size(1),
                                         _ptr(labels_u8), int(num_current), float(mins[0]), float(mins[1]),
                                         float(mins[2]), float(deltas[0]), float(deltas[1]), float(deltas[2]), X, Y,
                                         Z, int(num_classes), _ptr(ws), _ptr(vl), _ptr(pl), _stream())
    _lib.check(rc, "smos_vote_fused")
    _count(4)
    return vl, pl


def vote_stream(scans, current, crop_lo, crop_hi, mins, deltas, dims, num_classes=3):
    """One frame of streaming long-term voting over scans resident in HBM (smos_vote_stream).
    `scans`: list of (points (n, >=3) f32 CUDA, labels (n,) u8 CUDA, pose_diff 4x4 float64 array or None);
    `current`: index of the current scan (not transformed). crop_lo/crop_hi: float32 thresholds incl. eps.
    Returns (voxel_labels uint8 (X,Y,Z), point_labels int64 (n_current,))."""
    import numpy as np
    lib = _lib.load()
    descs = (_lib.VoteStreamScan * len(scans))()
    dev = scans[0][0].device
    total, rs = 0, None
    for d, (pts, lab, pose) in zip(descs, scans):
        _need_cuda(pts, "points")
        _need_f32(pts, "points")
        assert lab.dtype == torch.uint8 and lab.is_contiguous() and lab.is_cuda
        assert pts.dim() == 2 and pts.size(1) >= 3 and pts.stride(1) == 1
        r = pts.stride(0) if pts.size(0) > 1 else pts.size(1)
        rs = r if rs is None else rs
        assert r == rs, "all scans must share a row stride"
        d.points, d.labels, d.n = pts.data_ptr(), lab.data_ptr(), int(pts.size(0))
        if pose is None:
            d.transform = 0
        else:
            m = np.asarray(pose, dtype=np.float64).reshape(4, 4)
            d.transform = 1
            for k in range(12):
                d.pose_diff[k] = float(m[k // 4, k % 4])
        total += int(pts.size(0))
    X, Y, Z = (int(v) for v in dims)
    ws = _vote_ws(total, X, Y, Z, int(num_classes), dev)
    vl = torch.empty((X, Y, Z), dtype=torch.uint8, device=dev)
    pl = torch.empty((int(scans[current][0].size(0)),), dtype=torch.int64, device=dev)
    lo = (ctypes.c_float * 3)(*[float(v) for v in crop_lo])
    hi = (ctypes.c_float * 3)(*[float(v) for v in crop_hi])
    with torch.cuda.device(dev):
        rc = lib.smos_vote_stream(descs, len(scans), int(current), int(rs), lo, hi, float(mins[0]), float(mins[1]),
                                  float(mins[2]), float(deltas[0]), float(deltas[1]), float(deltas[2]), X, Y, Z,
                                  int(num_classes), _ptr(ws), _ptr(vl), _ptr(pl), _stream())
    _lib.check(rc, "smos_vote_stream")
    _count(4)
    return vl, pl


def memory_push(points, pred, cur_points, cur_pred, hist_points=None, hist_pred=None):
    """Long-term memory ring insert: (cur_points, cur_pred) -> (hist_points, hist_pred) if given, then
    (points (n, r) f32, pred (n,) u8) -> (cur_points, cur_pred). All CUDA, contiguous, same shapes. One kernel."""
    _need_cuda(points, "points")
    _need_f32(points, "points")
    ts = [points, pred, cur_points, cur_pred] + ([hist_points, hist_pred] if hist_points is not None else [])
    for t in ts:
        if not (t.is_cuda and t.is_contiguous()):
            raise RuntimeError("memory_push: tensors must be contiguous CUDA tensors")
    assert pred.dtype == torch.uint8 and cur_pred.dtype == torch.uint8 and cur_points.dtype == torch.float32
    assert cur_points.shape == points.shape and cur_pred.shape == pred.shape and points.dim() == 2
    if hist_points is not None:
        assert hist_points.shape == points.shape and hist_pred.shape == pred.shape
        assert hist_points.dtype == torch.float32 and hist_pred.dtype == torch.uint8
    with torch.cuda.device(points.device):
        rc = _lib.load().smos_memory_push(_ptr(points), _ptr(pred), int(points.size(0)), int(points.size(1)),
                                          _ptr(cur_points), _ptr(cur_pred), _ptr(hist_points), _ptr(hist_pred),
                                          _stream())
    _lib.check(rc, "smos_memory_push")
    _count(1)


def instance_vote_workspace(K, device):
    """Persistent accumulator for instance_vote(..., workspace=): zero now, left zero by every call."""
    n = int(_lib.load().smos_instance_vote_workspace_bytes(int(K)))
    return torch.zeros(n, dtype=torch.uint8, device=device)


def instance_vote(points, pred, box_lo, box_hi, count=None, workspace=None):
    """points (P, >=3) f32, pred (P,) int64, box_lo/box_hi (K, 3) f32 -> sums (K, 2) int64
    [static_sum, dynamic_sum] with dynamic points weighted 2. count: optional (1,) int32 CUDA tensor holding the
    number of valid boxes (rows beyond it stay zero) — read by the kernel, not by the host.
    workspace: instance_vote_workspace(K, device) owned by the calling stream — the vote then needs no zero fill of
    `sums` in front of it (one launch instead of two)."""
    _need_cuda(points, "points")
    _need_f32(points, "points")
    if pred.dtype != torch.int64:
        raise RuntimeError("pred must be int64")
    pred = pred.contiguous()
    box_lo = box_lo.to(torch.float32).contiguous()
    box_hi = box_hi.to(torch.float32).contiguous()
    assert points.dim() == 2 and points.size(1) >= 3 and points.stride(1) == 1
    P, K = int(points.size(0)), int(box_lo.size(0))
    if workspace is not None:
        if count is not None and (count.dtype != torch.int32 or not count.is_cuda):
            raise RuntimeError("count must be an int32 CUDA tensor")
        if workspace.numel() * workspace.element_size() < (2 * K + 2) * 8 or not workspace.is_cuda:
            raise RuntimeError("workspace too small (ops.instance_vote_workspace)")
        sums = torch.empty((K, 2), dtype=torch.int64, device=points.device)
        with torch.cuda.device(points.device):
            rc = _lib.load().smos_instance_vote_ws(_ptr(points), P, points.stride(0) if P > 1 else points.size(1),
                                                   _ptr(pred), _ptr(box_lo), _ptr(box_hi), K, _ptr(count), _ptr(workspace),
                                                   _ptr(sums), _stream())
        _lib.check(rc, "smos_instance_vote_ws")
        _count(1 if K else 0)
        return sums
    sums = torch.zeros((K, 2), dtype=torch.int64, device=points.device)
    with torch.cuda.device(points.device):
        if count is None:
            rc = _lib.load().smos_instance_vote(_ptr(points), P, points.stride(0) if P > 1 else points.size(1),
                                                _ptr(pred), _ptr(box_lo), _ptr(box_hi), K, _ptr(sums), _stream())
        else:
            if count.dtype != torch.int32 or not count.is_cuda:
                raise RuntimeError("count must be an int32 CUDA tensor")
            rc = _lib.load().smos_instance_vote_counted(_ptr(points), P,
                                                        points.stride(0) if P > 1 else points.size(1), _ptr(pred),
                                                        _ptr(box_lo), _ptr(box_hi), K, _ptr(count), _ptr(sums),
                                                        _stream())
    _lib.check(rc, "smos_instance_vote")
    _count(1)
    return sums


def cluster_boxes(points, pred_bf, eps=0.3, min_samples=5, min_cluster_points=30, z_lift=0.2):
    """Foreground selection + DBSCAN + kept-cluster boxes of cluster() (voxel_instance_voting.py:144-175).
    points (n, >=3) f32 CUDA, pred_bf (n,) integer CUDA. Returns a dict of device tensors: fg_index (n,) int32,
    fg_label (n,) int32, counts (3,) int32 [M, clusters, kept], box_lo / box_hi (Kcap, 3) f32, kept_label (Kcap,)
    int32, and the workspace smos_cluster_apply needs. No host synchronisation."""
    _need_cuda(points, "points")
    _need_f32(points, "points")
    assert points.dim() == 2 and points.size(1) >= 3 and points.stride(1) == 1
    n = int(points.size(0))
    if pred_bf.numel() != n:
        raise RuntimeError("pred_bf must have one entry per point")
    bf = pred_bf.reshape(-1).to(torch.int32).contiguous()
    dev = points.device
    lib = _lib.load()
    kcap = n // (int(min_cluster_points) + 1) + 1
    st = {
        "n": n,
        "fg_index": torch.empty((max(n, 1),), dtype=torch.int32, device=dev),
        "fg_label": torch.empty((max(n, 1),), dtype=torch.int32, device=dev),
        "counts": torch.empty((3,), dtype=torch.int32, device=dev),
        "box_lo": torch.empty((kcap, 3), dtype=torch.float32, device=dev),
        "box_hi": torch.empty((kcap, 3), dtype=torch.float32, device=dev),
        "kept_label": torch.empty((kcap,), dtype=torch.int32, device=dev),
        "workspace": torch.empty((int(lib.smos_cluster_workspace_bytes(n)),), dtype=torch.uint8, device=dev),
    }
    with torch.cuda.device(dev):
        rc = lib.smos_cluster_boxes(_ptr(points), n, points.stride(0) if n > 1 else points.size(1), _ptr(bf),
                                    float(eps), int(min_samples), int(min_cluster_points), float(z_lift),
                                    _ptr(st["workspace"]), _ptr(st["fg_index"]), _ptr(st["fg_label"]),
                                    _ptr(st["counts"]), _ptr(st["box_lo"]), _ptr(st["box_hi"]),
                                    _ptr(st["kept_label"]), _stream())
    _lib.check(rc, "smos_cluster_boxes")
    _count(9 if n else 0)
    return st


def cluster_apply(state, sums, pred):
    """Write-back of cluster() (voxel_instance_voting.py:184-191): pred (n,) int64 CUDA, updated in place."""
    if pred.dtype != torch.int64 or not pred.is_contiguous() or not pred.is_cuda:
        raise RuntimeError("pred must be a contiguous int64 CUDA tensor")
    if sums.dtype != torch.int64 or not sums.is_contiguous():
        raise RuntimeError("sums must be a contiguous int64 tensor")
    n = state["n"]
    if pred.numel() != n:
        raise RuntimeError("pred must have one entry per point")
    with torch.cuda.device(pred.device):
        rc = _lib.load().smos_cluster_apply(n, _ptr(state["workspace"]), _ptr(state["fg_index"]),
                                            _ptr(state["fg_label"]), _ptr(state["counts"]), _ptr(sums), _ptr(pred),
                                            _stream())
    _lib.check(rc, "smos_cluster_apply")
    _count(1 if n else 0)
    return pred
