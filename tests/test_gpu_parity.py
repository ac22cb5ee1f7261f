"""Parity of the sm_100a kernels (through the reference-shaped Python API -> ctypes -> C-ABI) against
(1) the committed outputs of the reference itself (tests/golden) and (2) the CPU oracle on seeded
synthetic inputs, including the BASELINE.json configuration sizes.

Bars (BASELINE.json north_star): voxel indices, scatter winners and vote counts BIT-EXACT; scatter-max
features bit-exact (max is exact); bilinear gathers and deformable attention within 1e-5 relative in
fp32 (plus a small atol at zero crossings)."""
import numpy as np
import pytest
import torch

from oracle import oracle as O

pytestmark = pytest.mark.gpu

RTOL, ATOL = 1e-5, 1e-6


def dev():
    return torch.device("cuda:0")


def t(a, dtype=None):
    x = torch.from_numpy(np.ascontiguousarray(a)).to(dev())
    return x.to(dtype) if dtype is not None else x


def synth_scan(rng, B, N, H, W, scale, n_valid=None):
    """BEV-like coordinates: valid points span the grid, ~1% in (-1,0) (truncate to cell 0), ~1% beyond the
    grid, pads at -1000 after n_valid (datasets/data_StreamMOS.py:570-571)."""
    n_valid = N if n_valid is None else n_valid
    c = np.stack([rng.uniform(0, H / scale[0], (B, N)), rng.uniform(0, W / scale[1], (B, N))], -1)
    k = max(1, N // 100)
    c[:, :k, 0] = rng.uniform(-0.999, 0, (B, k)) / scale[0]
    c[:, k:2 * k, 1] = W / scale[1] + rng.uniform(0, 2, (B, k))
    c[:, n_valid:] = -1000.0
    return c.astype(np.float32)[..., None]


# ------------------------------------------------------------------------------------------------
# VoxelMaxPool
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["pool_a", "pool_b", "pool_c"])
@pytest.mark.parametrize("layout", ["channel_major", "point_major"])
def test_pool_golden(golden, name, layout):
    from streammos_b200 import deep_point
    g = golden(name)
    feat = t(g["feat"])
    if layout == "point_major":
        feat = feat.contiguous(memory_format=torch.channels_last)
        assert feat.stride(1) == 1
    feat.requires_grad_(True)
    out = deep_point.VoxelMaxPool(feat, t(g["ind"]), (int(g["H"]), int(g["W"])), tuple(float(s) for s in g["scale"]))
    assert out.is_contiguous()
    assert np.array_equal(out.detach().cpu().numpy(), g["out"])
    out.backward(t(g["gout"]))
    assert np.array_equal(feat.grad.cpu().numpy(), g["gfeat"])  # scatter winners: every tied point


def test_pool_lower_boundary_voxel_max_idx(golden):
    """point_deep.cuda_kernel.* with the reference's 8/10-tensor signature (point_deep_cuda.cpp:22-62)."""
    from streammos_b200.point_deep import cuda_kernel
    g = golden("pool_a")
    feat, ind = t(g["feat"]), t(g["ind"])
    H, W = int(g["H"]), int(g["W"])
    B, C, N = feat.shape[:3]
    voxel_out = torch.zeros((B, C, H, W), device=dev())
    idx = torch.full((B, N), -1, dtype=torch.int64, device=dev())
    size_pt = torch.LongTensor([B, C, H, W]).to(dev())
    stride_pt = torch.LongTensor(list(voxel_out.stride())).to(dev())
    scale_pt = torch.FloatTensor(g["scale"]).to(dev())
    cuda_kernel.voxel_maxpooling_forward(feat, ind, voxel_out, idx, size_pt, stride_pt, size_pt[2:], scale_pt)
    o_ref, i_ref = O.voxel_maxpool_forward(g["feat"], g["ind"], (H, W), g["scale"], want_idx=True)
    assert np.array_equal(voxel_out.cpu().numpy(), o_ref)
    assert np.array_equal(idx.cpu().numpy(), i_ref)  # voxel indices bit-exact
    grad = torch.zeros_like(feat)
    cuda_kernel.voxel_maxpooling_backward(feat, ind, voxel_out, idx, grad, t(g["gout"]), size_pt, stride_pt,
                                          size_pt[2:], scale_pt)
    assert np.array_equal(grad.cpu().numpy(), g["gfeat"])
    with pytest.raises(RuntimeError):
        cuda_kernel.voxel_maxpooling_forward(feat.cpu(), ind, voxel_out, idx, size_pt, stride_pt, size_pt[2:],
                                             scale_pt)


POOL_CONFIG = [  # the five call sites per scan at N = 120k valid + pads
    (3, 64, (512, 512), (1.0, 1.0)),
    (1, 32, (32, 1024), (0.5, 0.5)),
    (1, 32, (256, 256), (0.5, 0.5)),
    (1, 64, (16, 512), (0.25, 0.25)),
    (1, 64, (128, 128), (0.25, 0.25)),
]


@pytest.mark.parametrize("B,C,size,scale", POOL_CONFIG)
@pytest.mark.parametrize("layout", ["channel_major", "point_major"])
def test_pool_config_sizes_vs_oracle(B, C, size, scale, layout):
    from streammos_b200 import deep_point
    rng = np.random.default_rng(7 + C + size[0])
    N, n_valid = 130000, 120000
    ind = synth_scan(rng, B, N, size[0], size[1], scale, n_valid)
    feat = rng.standard_normal((B, C, N, 1)).astype(np.float32)
    ref = O.voxel_maxpool_forward(feat, ind, size, scale)
    ft = t(feat)
    if layout == "point_major":
        ft = ft.contiguous(memory_format=torch.channels_last)
    ft.requires_grad_(True)
    out = deep_point.VoxelMaxPool(ft, t(ind), size, scale)
    assert np.array_equal(out.detach().cpu().numpy(), ref)
    gout = rng.standard_normal(ref.shape).astype(np.float32)
    out.backward(t(gout))
    gref = O.voxel_maxpool_backward(feat, ind, ref, gout, scale)
    assert np.array_equal(ft.grad.cpu().numpy().reshape(gref.shape), gref)


@pytest.mark.parametrize("B,C,N,size,scale", [
    (1, 1, 1, (4, 4), (1.0, 1.0)),       # single point
    (2, 3, 31, (5, 7), (1.0, 1.0)),      # ragged: N not a multiple of the warp, odd grid (scalar store path)
    (1, 7, 1000, (3, 1000), (1.0, 1.0)),  # wide, tile wider than tall
    (2, 40, 5000, (130, 70), (0.5, 0.5)),  # grid not a multiple of the tile, C not a power of two
    (1, 4, 2000, (8, 8), (1.0, 1.0)),    # heavy collisions: 2000 points into 64 cells
])
def test_pool_edge_shapes(B, C, N, size, scale):
    from streammos_b200 import deep_point
    rng = np.random.default_rng(N + C)
    ind = synth_scan(rng, B, N, size[0], size[1], scale)
    feat = rng.standard_normal((B, C, N, 1)).astype(np.float32)
    ref = O.voxel_maxpool_forward(feat, ind, size, scale)
    for fmt in (torch.contiguous_format, torch.channels_last):
        out = deep_point.VoxelMaxPool(t(feat).contiguous(memory_format=fmt), t(ind), size, scale)
        assert np.array_equal(out.cpu().numpy(), ref)


def test_pool_all_invalid_and_negative_maxima():
    from streammos_b200 import deep_point
    B, C, N = 1, 4, 500
    ind = np.full((B, N, 2, 1), -1000.0, np.float32)
    feat = -np.abs(np.random.default_rng(0).standard_normal((B, C, N, 1))).astype(np.float32) - 1
    out = deep_point.VoxelMaxPool(t(feat), t(ind), (16, 16), (1.0, 1.0))
    assert float(out.abs().max()) == 0.0  # empty grid is exactly zero
    ind[:, :, :, 0] = 3.5                 # everything into one cell, all features negative
    out = deep_point.VoxelMaxPool(t(feat), t(ind), (16, 16), (1.0, 1.0))
    ref = O.voxel_maxpool_forward(feat, ind, (16, 16), (1.0, 1.0))
    assert np.array_equal(out.cpu().numpy(), ref) and ref.min() < 0  # true (negative) max survives


def test_pool_permutation_invariant():
    """Size-independent property at the config size: permuting the points does not change the grid."""
    from streammos_b200 import deep_point
    rng = np.random.default_rng(11)
    B, C, N, size, scale = 1, 32, 120000, (256, 256), (0.5, 0.5)
    ind = synth_scan(rng, B, N, 512, 512, (1.0, 1.0))
    feat = rng.standard_normal((B, C, N, 1)).astype(np.float32)
    a = deep_point.VoxelMaxPool(t(feat), t(ind), size, scale)
    perm = rng.permutation(N)
    b = deep_point.VoxelMaxPool(t(feat[:, :, perm]), t(ind[:, perm]), size, scale)
    assert torch.equal(a, b)


def test_pool_batched_plans_match_single_plans():
    """ops.pool_plan_multi (four launches for all five pooling calls of a scan) gives the same grids."""
    from streammos_b200 import deep_point, ops
    rng = np.random.default_rng(3)
    N = 50000
    specs, feats = [], []
    for (B, C, size, scale) in POOL_CONFIG:
        ind = t(synth_scan(rng, B, N, size[0], size[1], scale, n_valid=47000))
        specs.append((ind, size, scale))
        feats.append(t(rng.standard_normal((B, C, N, 1)).astype(np.float32)))
    plans = ops.pool_plan_multi(specs)
    for (ind, size, scale), f, plan in zip(specs, feats, plans):
        a = deep_point.VoxelMaxPool(f, ind, size, scale, plan)
        b = deep_point.VoxelMaxPool(f, ind, size, scale)
        assert torch.equal(a, b)
        ref = O.voxel_maxpool_forward(f.cpu().numpy(), ind.cpu().numpy(), size, scale)
        assert np.array_equal(a.cpu().numpy(), ref)


@pytest.mark.parametrize("path", ["ldg", "tma"])
def test_pool_channel_major_permute_paths_random_shapes(path, monkeypatch):
    """Both permute kernels of the channel-major path (128-bit-load tiles and the tiled-TMA pipeline, selected with
    SMOS_PERM_LDG) over ragged shapes: N not a multiple of the 64-/128-point tiles, partial tiles, B > 1, channel
    counts around the 32-channel groups, dense duplicates (long runs) and all-invalid tails."""
    from streammos_b200 import deep_point
    monkeypatch.setenv("SMOS_PERM_LDG", "1" if path == "ldg" else "0")
    rng = np.random.default_rng(11 if path == "ldg" else 12)
    shapes = [(1, 64, 120000, (512, 512), (1.0, 1.0)), (3, 64, 4100, (64, 64), (1.0, 1.0)), (2, 32, 132, (8, 8), (0.5, 0.5)),
              (1, 96, 8192, (16, 16), (0.25, 0.25)), (2, 128, 2052, (32, 8), (1.0, 0.5)), (1, 8, 64, (4, 4), (1.0, 1.0)),
              (1, 64, 60, (4, 4), (1.0, 1.0)), (2, 160, 1028, (12, 20), (1.0, 1.0))]
    for B, C, N, size, scale in shapes:
        ind = synth_scan(rng, B, N, size[0], size[1], scale, n_valid=max(1, N - N // 7))
        ind[:, : N // 3] = np.floor(ind[:, : N // 3] / 3) * 3          # many equal neighbours: long merged runs
        feat = rng.standard_normal((B, C, N, 1)).astype(np.float32)
        out = deep_point.VoxelMaxPool(t(feat), t(ind), size, scale)
        ref = O.voxel_maxpool_forward(feat, ind, size, scale)
        assert np.array_equal(out.cpu().numpy(), ref), (path, B, C, N, size)


def test_pool_and_gather_val_loader_tta_shape():
    """The largest shapes the reference feeds the path: the val loader pads every frame to 160 000 points
    (config/StreamMOS.py:46) and test-time augmentation stacks 4 flips (datasets/data_StreamMOS.py:495-513), so
    VoxelMaxPool #1 sees B' = 4 x 3 = 12 batch entries. Forward bit-exact against the oracle, then a batched
    cell-order gather back (B = 4)."""
    from streammos_b200 import deep_point, ops
    rng = np.random.default_rng(160000)
    Bp, C, N, size = 12, 64, 160000, (512, 512)
    ind = synth_scan(rng, Bp, N, size[0], size[1], (1.0, 1.0), n_valid=121000)
    feat = np.maximum(rng.standard_normal((Bp, C, N, 1)).astype(np.float32), 0)
    out = deep_point.VoxelMaxPool(t(feat), t(ind), size, (1.0, 1.0))
    ref = O.voxel_maxpool_forward(feat, ind, size, (1.0, 1.0))
    assert out.shape == (Bp, C, 512, 512) and np.array_equal(out.cpu().numpy(), ref)
    del out
    B, Cg, H, W, scale = 4, 32, 256, 256, (0.5, 0.5)
    grid = rng.standard_normal((B, Cg, H, W)).astype(np.float32)
    co = ind[:B]
    plan = ops.pool_plan(t(co), (H, W), scale)
    got = ops.bilinear_gather_forward(t(grid), t(co), scale, True, order=plan)
    np.testing.assert_allclose(got[..., 0].cpu().numpy(), O.bilinear_sample(grid, co, scale), rtol=RTOL, atol=ATOL)


def test_pool_rejects_cpu_tensors():
    from streammos_b200 import deep_point
    with pytest.raises(RuntimeError):
        deep_point.VoxelMaxPool(torch.zeros(1, 2, 8, 1), torch.zeros(1, 8, 2, 1), (4, 4), (1.0, 1.0))


# ------------------------------------------------------------------------------------------------
# BilinearSample
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["bilinear_a", "bilinear_b"])
@pytest.mark.parametrize("grid_fmt", ["nchw", "nhwc"])
@pytest.mark.parametrize("point_major_out", [False, True])
def test_bilinear_golden(golden, name, grid_fmt, point_major_out):
    from streammos_b200.backbone import BilinearSample
    g = golden(name)
    grid = t(g["grid"])
    if grid_fmt == "nhwc":
        grid = grid.contiguous(memory_format=torch.channels_last)
    grid.requires_grad_(True)
    m = BilinearSample(in_dim=grid.shape[1], scale_rate=tuple(float(s) for s in g["scale"]))
    m.point_major_out = point_major_out
    out = m(grid, t(g["coord"]))
    assert out.shape == g["out"].shape
    if point_major_out:
        assert out.stride(1) == 1
    np.testing.assert_allclose(out.detach().cpu().numpy(), g["out"], rtol=RTOL, atol=ATOL)
    out.backward(t(g["gout"]))
    np.testing.assert_allclose(grid.grad.cpu().numpy(), g["ggrid"], rtol=1e-4, atol=1e-4)


BILINEAR_CONFIG = [  # the five gathers per scan
    (32, 256, 256, (0.5, 0.5)), (32, 32, 1024, (0.5, 0.5)), (64, 128, 128, (0.25, 0.25)),
    (64, 16, 512, (0.25, 0.25)), (64, 256, 256, (0.5, 0.5)),
]


@pytest.mark.parametrize("C,H,W,scale", BILINEAR_CONFIG)
def test_bilinear_config_sizes_vs_oracle(C, H, W, scale):
    from streammos_b200 import ops
    rng = np.random.default_rng(C + H)
    N = 120000
    coord = synth_scan(rng, 1, N, H, W, scale, n_valid=118000)
    grid = rng.standard_normal((1, C, H, W)).astype(np.float32)
    ref = O.bilinear_sample(grid, coord, scale)
    for fmt in (torch.contiguous_format, torch.channels_last):
        for pm in (False, True):
            out = ops.bilinear_gather_forward(t(grid).contiguous(memory_format=fmt), t(coord), scale, pm)
            np.testing.assert_allclose(out[..., 0].cpu().numpy(), ref, rtol=RTOL, atol=ATOL)
    # pads (coord -1000) sample nothing: exact zeros
    assert float(out[:, :, 118000:].abs().max()) == 0.0


@pytest.mark.parametrize("C,H,W,scale", [(32, 256, 256, (0.5, 0.5)), (64, 16, 512, (0.25, 0.25)), (5, 31, 47, (1.0, 1.0))])
def test_bilinear_cell_order_matches_scan_order(C, H, W, scale):
    """Visiting the points in the cell order of a pooling plan (any grid / scale built from the same
    coordinates) is a pure re-ordering: bit-identical to the scan-order kernel, out-of-grid pads included."""
    from streammos_b200 import ops
    rng = np.random.default_rng(C * H)
    B, N = 2, 60001
    coord = synth_scan(rng, B, N, H, W, scale, n_valid=57000)
    coord[1, 100:300] = 1e9                                     # far outside (int32 overflow range) on one batch
    grid = rng.standard_normal((B, C, H, W)).astype(np.float32)
    ref = O.bilinear_sample(grid, coord, scale)
    plans = [ops.pool_plan(t(coord), (H, W), scale), ops.pool_plan(t(coord), (7, 9), (0.01, 0.02))]
    plans += ops.pool_plan_multi([(t(coord), (H // 2 + 1, W // 2 + 1), (scale[0] / 2, scale[1] / 2))])
    for fmt in (torch.contiguous_format, torch.channels_last):
        g = t(grid).contiguous(memory_format=fmt)
        base = ops.bilinear_gather_forward(g, t(coord), scale, True)
        for plan in plans:
            out = ops.bilinear_gather_forward(g, t(coord), scale, True, order=plan)
            assert out.stride(1) == 1 and torch.equal(out, base)
        np.testing.assert_allclose(out[..., 0].cpu().numpy(), ref, rtol=RTOL, atol=ATOL)
    # sampling records emitted by the plan build (same grid and scale): bit-identical again, no coordinates read
    tp = ops.pool_plan_multi([(t(coord), (H, W), scale), (t(coord), (H + 3, W), scale)], gather_taps=[True, True])
    assert tp[0].taps is not None and tp[0].taps.numel() == B * N * 48
    for fmt in (torch.contiguous_format, torch.channels_last):
        g = t(grid).contiguous(memory_format=fmt)
        base = ops.bilinear_gather_forward(g, t(coord), scale, True)
        ops.reset_launch_count()
        assert torch.equal(ops.bilinear_gather_forward(g, t(coord), scale, True, order=tp[0]), base)
        # a plan of another geometry only lends its order, its records do not apply
        assert torch.equal(ops.bilinear_gather_forward(g, t(coord), scale, True, order=tp[1]), base)
    # channel-major outputs ignore the order (same kernel as without it)
    out = ops.bilinear_gather_forward(t(grid), t(coord), scale, False, order=plans[0])
    assert out.is_contiguous() and torch.equal(out, base.contiguous())
    with pytest.raises(RuntimeError):
        ops.bilinear_gather_forward(t(grid)[:1], t(coord)[:1], scale, True, order=plans[0])


@pytest.mark.parametrize("C,H,W,scale,N", [(64, 128, 128, (0.25, 0.25), 120000), (32, 256, 256, (0.5, 0.5), 60000),
                                          (32, 32, 1024, (0.5, 0.5), 60000), (3, 53, 47, (1.0, 1.0), 5000),
                                          (2, 3, 12500, (1.0, 1.0), 4000)])
def test_bilinear_backward_vs_oracle(C, H, W, scale, N):
    """Gradient wrt the grid against the float64-accumulating oracle: row bands in shared memory (several bands per
    plane, a ragged last band), and the global-atomics path for rows wider than a band (W = 12500)."""
    from streammos_b200 import ops
    rng = np.random.default_rng(C * H + W)
    B = 2 if C <= 3 else 1
    coord = np.concatenate([synth_scan(rng, 1, N, H, W, scale, n_valid=N - N // 50) for _ in range(B)])
    gout = rng.standard_normal((B, C, N, 1)).astype(np.float32)
    want = O.bilinear_sample_backward(gout, coord, scale, H, W)
    tol = 1e-5 * np.abs(want).max() + 1e-5
    got = ops.bilinear_gather_backward(t(gout), t(coord), scale, H, W)
    assert got.shape == (B, C, H, W) and bool(torch.isfinite(got).all())
    assert np.abs(got.cpu().numpy() - want).max() <= 4 * tol
    # point-major gradient rows (channels_last strides) take the same path
    gpm = t(gout).permute(0, 2, 1, 3).contiguous().permute(0, 2, 1, 3)
    assert gpm.stride(1) == 1
    got2 = ops.bilinear_gather_backward(gpm, t(coord), scale, H, W)
    assert np.abs(got2.cpu().numpy() - want).max() <= 4 * tol


def test_bilinear_multi_sample_dim_and_linearity():
    from streammos_b200 import ops
    rng = np.random.default_rng(5)
    B, C, H, W, N, S = 2, 8, 20, 30, 777, 3
    grid = rng.standard_normal((B, C, H, W)).astype(np.float32)
    coord = rng.uniform(-2, 32, (B, N, 2, S)).astype(np.float32)
    out = ops.bilinear_gather_forward(t(grid), t(coord), (1.0, 1.0))
    assert out.shape == (B, C, N, S)
    for s in range(S):
        ref = O.bilinear_sample(grid, coord[..., s], (1.0, 1.0))
        np.testing.assert_allclose(out[..., s].cpu().numpy(), ref, rtol=RTOL, atol=ATOL)
    # linearity in the grid: gather(2a + b) == 2 gather(a) + gather(b)
    g2 = rng.standard_normal((B, C, H, W)).astype(np.float32)
    lhs = ops.bilinear_gather_forward(t(2 * grid + g2), t(coord), (1.0, 1.0))
    rhs = 2 * out + ops.bilinear_gather_forward(t(g2), t(coord), (1.0, 1.0))
    torch.testing.assert_close(lhs, rhs, rtol=1e-4, atol=1e-5)


# ------------------------------------------------------------------------------------------------
# MSDeformAttn
# ------------------------------------------------------------------------------------------------
MSDA = ["msda_reftest", "msda_reftest_d30", "msda_reftest_d32", "msda_reftest_d71", "msda_streammos_small"]


@pytest.mark.parametrize("name", MSDA)
@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
def test_msda_golden(golden, name, dtype):
    from streammos_b200.functions import MSDeformAttnFunction
    g = golden(name)
    value = t(g["value"], dtype).requires_grad_(True)
    loc = t(g["loc"], dtype).requires_grad_(True)
    attn = t(g["attn"], dtype).requires_grad_(True)
    out = MSDeformAttnFunction.apply(value, t(g["shapes"]), t(g["lsi"]), loc, attn, 2)
    out.backward(t(g["gout"], dtype))
    scale = np.abs(g["out64"]).max()
    if dtype == torch.float64:  # deformattn/test.py:41: allclose defaults in double
        tol = dict(rtol=1e-7, atol=1e-10)
        gtol = dict(rtol=1e-6, atol=1e-9)
    else:                       # north_star: 1e-5 relative in fp32
        tol = dict(rtol=RTOL, atol=1e-6 * max(scale, 1e-3))
        # fp32 gradients against the fp64 reference: rtol 1e-4 plus 1e-5 of each gradient's own scale. d/d(loc) is
        # a difference of neighbouring pixels times the map size (cancellation of O(|value|) terms) and grad_value is
        # accumulated with order-dependent fp32 atomics, so a pure 1e-5 relative bound is not attainable in fp32 —
        # the reference's own float check uses rtol 1e-2 / atol 1e-3 (deformattn/test.py:56)
        gtol = None
    np.testing.assert_allclose(out.detach().cpu().numpy(), g["out64"], **tol)
    for got, want in ((value.grad, g["gvalue"]), (attn.grad, g["gattn"]), (loc.grad, g["gloc"])):
        k = gtol if gtol is not None else dict(rtol=1e-4, atol=1e-5 * max(np.abs(want).max(), 1e-6))
        np.testing.assert_allclose(got.cpu().numpy(), want, **k)


@pytest.mark.parametrize("name", ["msda_module_streammos", "msda_module_boxes"])
@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
def test_msda_fused_forward_golden(golden, name, dtype):
    """One kernel for softmax + sampling locations + sampling core against what the reference module
    computed on the CPU (its own forward, core = ms_deform_attn_core_pytorch), 2-d and 4-d reference points."""
    from streammos_b200 import ops
    g = golden(name)
    out = ops.ms_deform_attn_fused_forward(t(g["value"], dtype), t(g["shapes"]), t(g["lsi"]), t(g["offsets"], dtype),
                                           t(g["logits"], dtype), t(g["ref"], dtype))
    ref64 = O.ms_deform_attn_fused_forward(g["value"], g["shapes"], g["lsi"], g["offsets"], g["logits"], g["ref"])
    scale = np.abs(ref64).max()
    tol = dict(rtol=1e-7, atol=1e-10) if dtype == torch.float64 else dict(rtol=RTOL, atol=2e-6 * scale)
    np.testing.assert_allclose(out.cpu().numpy(), ref64, **tol)
    np.testing.assert_allclose(out.cpu().numpy(), g["core_out"], rtol=1e-4, atol=1e-5 * scale)
    # and the unfused core on the module's own locations / weights agrees with the fused kernel
    un = ops.ms_deform_attn_forward(t(g["value"], dtype), t(g["shapes"]), t(g["lsi"]), t(g["loc"], dtype),
                                    t(g["attn"], dtype))
    np.testing.assert_allclose(out.cpu().numpy(), un.cpu().numpy(), rtol=1e-4, atol=1e-5 * scale)


@pytest.mark.parametrize("name", ["msda_module_streammos", "msda_module_boxes"])
def test_msdeformattn_module_golden(golden, name):
    """streammos_b200.modules.MSDeformAttn (reference constructor and parameter names) loaded with the reference
    module's weights: inference (fused kernel) and autograd (reference sequence) paths against the reference output."""
    from streammos_b200.modules import MSDeformAttn
    g = golden(name)
    d_model, L, M, P = (int(v) for v in g["dims"])
    m = MSDeformAttn(d_model, L, M, P)
    sd = {k[2:].replace("__", "."): torch.from_numpy(np.asarray(g[k])) for k in list(g.keys()) if k.startswith("w_")}
    m.load_state_dict(sd, strict=True)
    m = m.cuda().eval()
    args = (t(g["query"]), t(g["ref"]), t(g["src"]), t(g["shapes"]), t(g["lsi"]))
    with torch.no_grad():
        out = m(*args)
    np.testing.assert_allclose(out.cpu().numpy(), g["out"], rtol=1e-4, atol=2e-5)
    q = args[0].clone().requires_grad_(True)
    out2 = m(q, *args[1:])
    np.testing.assert_allclose(out2.detach().cpu().numpy(), g["out"], rtol=1e-4, atol=2e-5)
    out2.sum().backward()
    assert torch.isfinite(q.grad).all() and float(q.grad.abs().sum()) > 0
    with pytest.raises(ValueError):
        m(args[0], args[1][..., :1].repeat(1, 1, 1, 3), *args[2:])


def test_msda_forward_many_samples_and_wide_heads():
    """Shapes that exercise the sample batching of the forward kernel: L*P not a multiple of the batch or of the lane
    group (3 levels x 3 points = 9 samples), D = 8 (two lanes per group), D = 160 (more channel quads than lanes),
    and samples far outside the maps."""
    from streammos_b200 import ops
    rng = np.random.default_rng(77)
    for D, M in ((8, 3), (160, 2), (5, 2)):
        shapes = np.array([[7, 5], [4, 6], [3, 3]], np.int64)
        lsi = np.concatenate(([0], np.cumsum(shapes.prod(1))[:-1])).astype(np.int64)
        S, B, Q, L, P = int(shapes.prod(1).sum()), 2, 37, 3, 3
        value = rng.standard_normal((B, S, M, D))
        loc = rng.uniform(-0.4, 1.4, (B, Q, M, L, P, 2))
        attn = rng.uniform(0, 1, (B, Q, M, L, P))
        ref = O.ms_deform_attn_forward(value, shapes, lsi, loc, attn)
        for dtype, tol in ((torch.float64, dict(rtol=1e-7, atol=1e-10)), (torch.float32, dict(rtol=RTOL, atol=1e-5))):
            out = ops.ms_deform_attn_forward(t(value, dtype), t(shapes), t(lsi), t(loc, dtype), t(attn, dtype))
            np.testing.assert_allclose(out.cpu().numpy(), ref, **tol)


def _streammos_msda_inputs(rng, B, Hs=64, Ws=64, M=4, D=32, P=4):
    value = rng.standard_normal((B, Hs * Ws, M, D)).astype(np.float32)
    ys, xs = np.meshgrid(np.linspace(0.5, Hs - 0.5, Hs), np.linspace(0.5, Ws - 0.5, Ws), indexing="ij")
    ref_pts = np.stack((xs.reshape(-1) / Ws, ys.reshape(-1) / Hs), -1)
    loc = ref_pts[None, :, None, None, None, :] + rng.standard_normal((B, Hs * Ws, M, 1, P, 2)) * 3.0 / Hs
    a = rng.standard_normal((B, Hs * Ws, M, P))
    attn = np.exp(a) / np.exp(a).sum(-1, keepdims=True)
    shapes = np.array([[Hs, Ws]], np.int64)
    lsi = np.array([0], np.int64)
    return value, shapes, lsi, loc.astype(np.float32), attn.reshape(B, Hs * Ws, M, 1, P).astype(np.float32)


@pytest.mark.parametrize("B", [1, 4])
def test_msda_config_shape_vs_oracle(B):
    """config #3: value (B,4096,4,32), [[64,64]], 4 points — fwd and bwd against the fp64 oracle."""
    from streammos_b200 import MultiScaleDeformableAttention as MSDA_mod
    rng = np.random.default_rng(100 + B)
    value, shapes, lsi, loc, attn = _streammos_msda_inputs(rng, B)
    ref = O.ms_deform_attn_forward(value, shapes, lsi, loc, attn)
    out = MSDA_mod.ms_deform_attn_forward(t(value), t(shapes), t(lsi), t(loc), t(attn), 256)
    assert out.shape == (B, 4096, 128)
    np.testing.assert_allclose(out.cpu().numpy(), ref, rtol=RTOL, atol=1e-6 * np.abs(ref).max())
    gout = rng.standard_normal(ref.shape).astype(np.float32)
    gv, gl, ga = MSDA_mod.ms_deform_attn_backward(t(value), t(shapes), t(lsi), t(loc), t(attn), t(gout), 256)
    rv, rl, ra = O.ms_deform_attn_backward(value, shapes, lsi, loc, attn, gout)
    np.testing.assert_allclose(gv.cpu().numpy(), rv, rtol=1e-4, atol=1e-5 * np.abs(rv).max())
    np.testing.assert_allclose(ga.cpu().numpy(), ra, rtol=1e-4, atol=1e-5 * np.abs(ra).max())
    np.testing.assert_allclose(gl.cpu().numpy(), rl, rtol=1e-4, atol=1e-5 * np.abs(rl).max())


@pytest.mark.parametrize("channels", [30, 32, 64, 71])
def test_msda_gradcheck_double(channels):
    """deformattn/test.py:63-78 check_gradient_numerical, same shapes and seed."""
    from torch.autograd import gradcheck
    from streammos_b200.functions import MSDeformAttnFunction
    torch.manual_seed(3)
    N, M, Lq, L, P = 1, 2, 2, 2, 2
    shapes = torch.as_tensor([(6, 4), (3, 2)], dtype=torch.long, device=dev())
    lsi = torch.cat((shapes.new_zeros((1,)), shapes.prod(1).cumsum(0)[:-1]))
    S = int(shapes.prod(1).sum())
    value = (torch.rand(N, S, M, channels, device=dev()) * 0.01).double().requires_grad_(True)
    loc = torch.rand(N, Lq, M, L, P, 2, device=dev()).double().requires_grad_(True)
    attn = torch.rand(N, Lq, M, L, P, device=dev()) + 1e-5
    attn = (attn / attn.sum(-1, keepdim=True).sum(-2, keepdim=True)).double().requires_grad_(True)
    assert gradcheck(MSDeformAttnFunction.apply, (value, shapes, lsi, loc, attn, 2))


@pytest.mark.parametrize("channels", [1025, 2048, 3096])
def test_msda_gradcheck_double_large_channels(channels):
    """The large channel counts of deformattn/test.py:85 (the reference routes D > 1024 to its multi-block reduction
    kernels, ms_deform_im2col_cuda.cuh:1015-1130). Same shapes and seed; gradcheck in fast mode (random projections of
    the Jacobian) because the full numerical Jacobian of a 186 k-element value tensor is 18 GB."""
    from torch.autograd import gradcheck
    from streammos_b200.functions import MSDeformAttnFunction
    torch.manual_seed(3)
    N, M, Lq, L, P = 1, 2, 2, 2, 2
    shapes = torch.as_tensor([(6, 4), (3, 2)], dtype=torch.long, device=dev())
    lsi = torch.cat((shapes.new_zeros((1,)), shapes.prod(1).cumsum(0)[:-1]))
    S = int(shapes.prod(1).sum())
    value = (torch.rand(N, S, M, channels, device=dev()) * 0.01).double().requires_grad_(True)
    loc = torch.rand(N, Lq, M, L, P, 2, device=dev()).double().requires_grad_(True)
    attn = torch.rand(N, Lq, M, L, P, device=dev()) + 1e-5
    attn = (attn / attn.sum(-1, keepdim=True).sum(-2, keepdim=True)).double().requires_grad_(True)
    # grad_value is accumulated with atomics (as in the reference): the summation order, hence the last bits, vary
    assert gradcheck(MSDeformAttnFunction.apply, (value, shapes, lsi, loc, attn, 2), fast_mode=True, nondet_tol=1e-10)
    # and the analytical gradients equal torch autograd through the fp64 oracle formula, element for element
    out = MSDeformAttnFunction.apply(value, shapes, lsi, loc, attn, 2)
    gout = torch.rand_like(out)
    out.backward(gout)
    rv, rl, ra = O.ms_deform_attn_backward(value.detach().cpu().numpy(), shapes.cpu().numpy(), lsi.cpu().numpy(),
                                           loc.detach().cpu().numpy(), attn.detach().cpu().numpy(), gout.cpu().numpy())
    np.testing.assert_allclose(value.grad.cpu().numpy(), rv, rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(loc.grad.cpu().numpy(), rl, rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(attn.grad.cpu().numpy(), ra, rtol=1e-9, atol=1e-12)


def test_msda_forward_into_caller_buffer():
    from streammos_b200 import ops
    rng = np.random.default_rng(5)
    value = rng.standard_normal((1, 64 * 64, 4, 32)).astype(np.float32)
    loc = rng.uniform(0, 1, (1, 4096, 4, 1, 4, 2)).astype(np.float32)
    attn = rng.uniform(0, 1, (1, 4096, 4, 1, 4)).astype(np.float32)
    shapes, lsi = np.array([[64, 64]], np.int64), np.array([0], np.int64)
    args = [t(a) for a in (value, shapes, lsi, loc, attn)]
    want = ops.ms_deform_attn_forward(*args)
    buf = torch.full((1, 4096, 128), 7.0, device=dev())
    got = ops.ms_deform_attn_forward(*args, out=buf)
    assert got.data_ptr() == buf.data_ptr() and torch.equal(got, want)
    with pytest.raises(RuntimeError):
        ops.ms_deform_attn_forward(*args, out=args[0].view(1, 4096, 128))


def test_msda_error_behaviour():
    from streammos_b200 import MultiScaleDeformableAttention as MSDA_mod
    rng = np.random.default_rng(0)
    value, shapes, lsi, loc, attn = _streammos_msda_inputs(rng, 1, 8, 8)
    args = [t(value), t(shapes), t(lsi), t(loc), t(attn)]
    with pytest.raises(RuntimeError, match="contiguous"):
        MSDA_mod.ms_deform_attn_forward(args[0].transpose(2, 3), *args[1:], 256)
    with pytest.raises(RuntimeError, match="CUDA"):
        MSDA_mod.ms_deform_attn_forward(args[0].cpu(), *args[1:], 256)
    with pytest.raises(RuntimeError):
        MSDA_mod.ms_deform_attn_forward(args[0].half(), *args[1:], 256)


# ------------------------------------------------------------------------------------------------
# Voting
# ------------------------------------------------------------------------------------------------
def test_voting_golden(golden):
    from streammos_b200 import voting
    g = golden("voting_a")
    size = tuple(int(s) for s in g["size"])
    rx, ry, rz = tuple(g["rx"]), tuple(g["ry"]), tuple(g["rz"])
    q = voting.Quantize(t(g["pts"]), range_x=rx, range_y=ry, range_z=rz, size=size)
    assert np.array_equal(q.cpu().numpy(), g["quan"])
    coords = q.to(torch.int64)
    assert np.array_equal(coords.cpu().numpy(), g["coords"])          # voxel indices bit-exact
    vl = voting.determine_voxel_labels(coords, t(g["labels"]), size)  # num_classes via labels.max() like the ref
    assert vl.dtype == torch.int64 and np.array_equal(vl.cpu().numpy(), g["voxel_labels"])
    vl3 = voting.determine_voxel_labels(coords, t(g["labels"]), size, num_classes=3)
    assert torch.equal(vl, vl3)
    pl = voting.get_point_labels_from_voxel_labels(t(g["cur"]), vl, size)
    assert np.array_equal(pl.cpu().numpy(), g["point_labels"])


@pytest.mark.parametrize("num_classes", [3, 5])
def test_voting_config_size_vs_oracle(num_classes):
    """config #1 shape: 8 history + 1 current scans x 120k points into 512 x 512 x 30."""
    from streammos_b200 import ops, voting
    rng = np.random.default_rng(42 + num_classes)
    P, Pc = 9 * 120000, 120000
    size = (512, 512, 30)
    rx, ry, rz = (-50.0, 50.0), (-50.0, 50.0), (-4.0, 2.0)
    r = np.abs(rng.standard_normal(P)) * 18.0
    th = rng.uniform(0, 2 * np.pi, P)
    pts = np.stack([np.clip(r * np.cos(th), -49.99, 49.99), np.clip(r * np.sin(th), -49.99, 49.99),
                    np.clip(rng.normal(-1.5, 0.6, P), -3.99, 1.99), rng.uniform(0, 1, P)], -1).astype(np.float32)
    labels = rng.integers(0, num_classes, P).astype(np.int64)
    q_ref = O.quantize(pts, rx, ry, rz, size)
    q = voting.Quantize(t(pts), rx, ry, rz, size)
    assert np.array_equal(q.cpu().numpy(), q_ref)
    coords = q.to(torch.int64)
    vl_ref = O.determine_voxel_labels(q_ref.astype(np.int64), labels, size, num_classes)
    vl = voting.determine_voxel_labels(coords, t(labels), size, num_classes=num_classes)
    assert np.array_equal(vl.cpu().numpy(), vl_ref)
    pl = voting.get_point_labels_from_voxel_labels(coords[P - Pc:], vl, size)
    pl_ref = O.get_point_labels_from_voxel_labels(q_ref.astype(np.int64)[P - Pc:], vl_ref, size)
    assert np.array_equal(pl.cpu().numpy(), pl_ref)
    # fused streaming variant: same answers from float xyz + uint8 labels
    d = [np.float32((rr[1] - rr[0]) / s) for rr, s in zip((rx, ry, rz), size)]
    vl8, plf = ops.vote_fused(t(pts), t(labels.astype(np.uint8)), Pc, (rx[0], ry[0], rz[0]), d, size, num_classes)
    assert np.array_equal(vl8.cpu().numpy().astype(np.int64), vl_ref)
    assert np.array_equal(plf.cpu().numpy(), pl_ref)
    # checksum of checksums: votes are conserved — every in-range point lands in exactly one voxel
    assert int((vl > 0).sum()) <= P


def test_voting_empty_and_out_of_range():
    from streammos_b200 import voting
    size = (8, 8, 4)
    coords = torch.tensor([[0, 0, 0], [0, 0, 0], [7, 7, 3], [9, 0, 0], [-1, 2, 2]], dtype=torch.int64, device=dev())
    labels = torch.tensor([2, 2, 1, 2, 2], dtype=torch.int64, device=dev())
    vl = voting.determine_voxel_labels(coords, labels, size, num_classes=3)
    assert int(vl[0, 0, 0]) == 2 and int(vl[7, 7, 3]) == 1 and int(vl.sum()) == 3
    pl = voting.get_point_labels_from_voxel_labels(coords, vl, size)
    assert pl.tolist() == [2, 2, 1, 0, 0]
    # tie -> lowest class
    vl = voting.determine_voxel_labels(coords[:2], torch.tensor([2, 1], device=dev()), size, num_classes=3)
    assert int(vl[0, 0, 0]) == 1


@pytest.mark.parametrize("size,P", [((3, 3, 3), 500), ((5, 7, 3), 4000), ((1, 1, 1), 10)])
def test_voting_odd_grid_sizes_in_place_counters(size, P):
    """The int64 label slots double as the packed vote counters: odd voxel counts exercise the scalar tail of the
    in-place argmax pass, and the result must not depend on what the output buffer held before."""
    from streammos_b200 import ops
    rng = np.random.default_rng(P)
    coords = np.stack([rng.integers(-1, s + 1, P) for s in size], -1).astype(np.int64)   # some out of range
    labels = rng.integers(0, 3, P).astype(np.int64)
    ref = O.determine_voxel_labels(coords, labels, size, 3)
    for _ in range(2):
        torch.full((64,), 7, dtype=torch.int64, device=dev())                          # dirty the allocator's cache
        vl = ops.vote_voxel_labels(t(coords), t(labels), size, 3)
        assert np.array_equal(vl.cpu().numpy(), ref)


@pytest.mark.parametrize("n", [1, 17, 4099, 120000])
def test_memory_push_moves_current_into_history(n):
    from streammos_b200 import ops
    rng = np.random.default_rng(n)
    new_p, new_l = rng.standard_normal((n, 4)).astype(np.float32), rng.integers(0, 3, n).astype(np.uint8)
    cur_p, cur_l = rng.standard_normal((n, 4)).astype(np.float32), rng.integers(0, 3, n).astype(np.uint8)
    ring_p, ring_l = torch.zeros(3, n, 4, device=dev()), torch.zeros(3, n, dtype=torch.uint8, device=dev())
    ring_p[2].copy_(t(cur_p)); ring_l[2].copy_(t(cur_l))
    ops.memory_push(t(new_p), t(new_l), ring_p[2], ring_l[2], ring_p[1], ring_l[1])
    assert np.array_equal(ring_p[1].cpu().numpy(), cur_p) and np.array_equal(ring_l[1].cpu().numpy(), cur_l)
    assert np.array_equal(ring_p[2].cpu().numpy(), new_p) and np.array_equal(ring_l[2].cpu().numpy(), new_l)
    assert float(ring_p[0].abs().sum()) == 0.0
    ops.memory_push(t(cur_p), t(cur_l), ring_p[2], ring_l[2])                            # no history slot: overwrite only
    assert np.array_equal(ring_p[2].cpu().numpy(), cur_p) and np.array_equal(ring_p[1].cpu().numpy(), cur_p)


@pytest.mark.parametrize("S,n,push", [(9, 120000, True), (9, 4099, True), (3, 17, True), (2, 1, False), (4, 50000, False)])
def test_vote_stage_equals_quantize_and_casts(S, n, push):
    """smos_vote_stage = ring insert + Quantize + the script's two .to(int64) casts (voxel_voting.py:234-241): bit-exact
    against the oracle's Quantize, numpy truncation and a plain ring update; odd n exercises the scalar stores."""
    from streammos_b200 import voting
    rng = np.random.default_rng(S * 1000 + n)
    ring_p = np.concatenate([rng.uniform(-52, 52, (S, n, 2)), rng.uniform(-4.5, 2.5, (S, n, 1)), rng.uniform(0, 1, (S, n, 1))],
                            -1).astype(np.float32)
    ring_p[:, ::11] = -1000.0                                       # pads
    ring_p[:, 1::13, 0] = np.float32(-50.0) - np.float32(1e-3)      # quantises into (-1, 0): truncates to 0
    ring_l = rng.integers(0, 3, (S, n)).astype(np.uint8)
    new_p = ring_p[0][::-1].copy() * np.float32(0.9)
    new_l = rng.integers(0, 3, n).astype(np.uint8)
    rx, ry, rz, size = (-50.0, 50.0), (-50.0, 50.0), (-4.0, 2.0), (512, 512, 30)
    dp, dl = t(ring_p), t(ring_l)
    cur, hist = S - 1, (S - 2 if S > 2 else -1)
    want_p, want_l = ring_p.copy(), ring_l.copy()
    if push:
        if hist >= 0:
            want_p[hist], want_l[hist] = ring_p[cur], ring_l[cur]
        want_p[cur], want_l[cur] = new_p, new_l
        q, coords, labels = voting.quantize_staged(dp, dl, rx, ry, rz, size, new_points=t(new_p), new_pred=t(new_l),
                                                   cur_slot=cur, hist_slot=hist)
    else:
        q, coords, labels = voting.quantize_staged(dp, dl, rx, ry, rz, size)
    assert np.array_equal(dp.cpu().numpy(), want_p) and np.array_equal(dl.cpu().numpy(), want_l)
    want_q = O.quantize(want_p.reshape(-1, 4), rx, ry, rz, size)
    assert np.array_equal(q.cpu().numpy(), want_q)
    assert np.array_equal(coords.cpu().numpy(), want_q.astype(np.int64))      # numpy / torch casts truncate
    assert np.array_equal(labels.cpu().numpy(), want_l.reshape(-1).astype(np.int64))
    # and the separate reference-shaped calls agree
    q2 = voting.Quantize(dp.view(-1, 4), rx, ry, rz, size)
    assert torch.equal(q2, q) and torch.equal(q2.to(torch.int64), coords)
    # want_q=False (what the stream harness uses: the float tensor is a dead temporary in the script): same casts
    q3, coords3, labels3 = voting.quantize_staged(dp, dl, rx, ry, rz, size, want_q=False)
    assert q3 is None and torch.equal(coords3, coords) and torch.equal(labels3, labels)
    # the script's Crop(fov -/+ 1e-4) in front of Quantize (voxel_voting.py:225-231): points outside the open box
    # (float32 thresholds, utils/transforms.py:155-157) take no part — coords -1 — all others are untouched
    _, coords4, labels4 = voting.quantize_staged(dp, dl, rx, ry, rz, size, want_q=False, crop_eps=1e-4)
    pts4 = want_p.reshape(-1, 4)
    lo = np.array([np.float32(r[0] + 1e-4) for r in (rx, ry, rz)], np.float32)
    hi = np.array([np.float32(r[1] - 1e-4) for r in (rx, ry, rz)], np.float32)
    inside = ((pts4[:, :3] > lo) & (pts4[:, :3] < hi)).all(1)
    assert pts4.shape[0] < 100 or (inside.any() and (~inside).any())
    want4 = np.where(inside[:, None], want_q.astype(np.int64), -1)
    assert np.array_equal(coords4.cpu().numpy(), want4) and torch.equal(labels4, labels)


def test_instance_vote_workspace_variant_needs_no_zero_fill():
    from streammos_b200 import ops
    rng = np.random.default_rng(19)
    P, K = 200000, 40
    pts = np.concatenate([rng.uniform(-50, 50, (P, 2)), rng.uniform(-4, 2, (P, 1)), rng.uniform(0, 1, (P, 1))],
                         1).astype(np.float32)
    pred = rng.integers(0, 3, P).astype(np.int64)
    c = np.concatenate([rng.uniform(-45, 45, (K, 2)), rng.uniform(-3, 1, (K, 1))], 1)
    half = rng.uniform(0.5, 4, (K, 3))
    lo, hi = (c - half).astype(np.float32), (c + half).astype(np.float32)
    want = O.instance_vote(pts, pred, lo, hi)
    ws = ops.instance_vote_workspace(K, dev())
    for rep in range(3):  # the workspace is left zero: repeated calls give the same totals
        got = ops.instance_vote(t(pts), t(pred), t(lo), t(hi), workspace=ws)
        assert np.array_equal(got.cpu().numpy(), want), rep
        assert not ws.cpu().numpy().any()
    for k in (K, 7, 0):
        count = torch.tensor([k], dtype=torch.int32, device=dev())
        got = ops.instance_vote(t(pts), t(pred), t(lo), t(hi), count=count, workspace=ws).cpu().numpy()
        assert np.array_equal(got[:k], want[:k]) and not got[k:].any(), k
    # more boxes than one shared-memory chunk
    K2 = 300
    c = np.concatenate([rng.uniform(-45, 45, (K2, 2)), rng.uniform(-3, 1, (K2, 1))], 1)
    half = rng.uniform(0.5, 4, (K2, 3))
    lo, hi = (c - half).astype(np.float32), (c + half).astype(np.float32)
    ws2 = ops.instance_vote_workspace(K2, dev())
    got = ops.instance_vote(t(pts), t(pred), t(lo), t(hi), workspace=ws2)
    assert np.array_equal(got.cpu().numpy(), O.instance_vote(pts, pred, lo, hi))


def test_voting_single_class2_vote_and_heavy_voxels():
    """Counter words of the in-place int64 path: one vote for class 2 is the word 2 (= its own label), several points of
    one voxel convert it concurrently, and ties resolve to the lowest class."""
    from streammos_b200 import ops
    size = (4, 4, 2)
    coords = np.array([[0, 0, 0]] * 1 + [[1, 1, 1]] * 7 + [[2, 2, 0]] * 4000 + [[3, 3, 1]] * 6, np.int64)
    labels = np.array([2] + [2] * 3 + [1] * 4 + list(np.arange(4000) % 3) + [0, 0, 1, 1, 2, 2], np.int64)
    ref = O.determine_voxel_labels(coords, labels, size, 3)
    assert ref[0, 0, 0] == 2 and ref[1, 1, 1] == 1 and ref[3, 3, 1] == 0
    perm = np.random.default_rng(0).permutation(len(labels))
    for order in (np.arange(len(labels)), perm):
        vl = ops.vote_voxel_labels(t(coords[order]), t(labels[order]), size, 3)
        assert np.array_equal(vl.cpu().numpy(), ref)


def test_instance_vote_golden(golden):
    from streammos_b200 import voting
    g = golden("instance_a")
    stat, dyn, label = voting.instance_vote_counts(t(g["local_pts"]), t(g["local_pred"]), t(g["corners"]))
    assert np.array_equal(stat.cpu().numpy(), g["stat"])
    assert np.array_equal(dyn.cpu().numpy(), g["dyn"])
    assert np.array_equal(label.cpu().numpy(), g["label"])


def test_cluster_golden(golden):
    """cluster() of voxel_instance_voting.py:144-193 on the device against the reference's own run (sklearn DBSCAN +
    scipy hull / Delaunay): DBSCAN labels, and the relabelled prediction, bit-exact."""
    from streammos_b200 import voting
    g = golden("cluster_a")
    fg = np.where(g["cur_bf"] == 2)[0]
    labels = voting.dbscan_fit_predict(t(g["cur_pts"][fg]))
    assert labels.dtype == torch.int32
    assert np.array_equal(labels.cpu().numpy(), g["fg_labels"])
    for name in ("cluster_a", "instance_a"):
        g = golden(name)
        pred = t(g["cur_pred"])
        out = voting.cluster(t(g["cur_pts"]), pred, t(g["cur_bf"].astype(np.int64)), t(g["local_pts"]),
                             t(g["local_pred"]))
        assert out.data_ptr() == pred.data_ptr()                 # in place, like the reference
        assert np.array_equal(out.cpu().numpy(), g["cluster_out"]), name


def _blobs(rng, n_blobs, n_noise, spread=8.0, lo=5, hi=400):
    parts = [rng.uniform(-spread, spread, 3) + rng.uniform(-1, 1, (int(rng.integers(lo, hi)), 3)) *
             rng.uniform(0.2, 1.4, 3) for _ in range(n_blobs)]
    parts.append(rng.uniform(-spread - 2, spread + 2, (n_noise, 3)))
    x = np.concatenate(parts).astype(np.float32)
    return x[rng.permutation(len(x))]


@pytest.mark.parametrize("seed,n_blobs,n_noise", [(0, 1, 0), (1, 4, 300), (2, 9, 1500), (3, 0, 777), (4, 12, 40),
                                                  (5, 30, 3000)])
def test_dbscan_vs_oracle(seed, n_blobs, n_noise):
    """Labels identical to the sequential scikit-learn algorithm (oracle): core points, cluster numbering by first
    core point, border points to the lowest-numbered adjacent cluster, noise."""
    from streammos_b200 import voting
    x = _blobs(np.random.default_rng(seed), n_blobs, n_noise)
    got = voting.dbscan_fit_predict(t(x)).cpu().numpy()
    want = O.dbscan(x)
    assert np.array_equal(got, want)


def test_dbscan_edges():
    from streammos_b200 import ops, voting
    z = torch.zeros((7, 3), device=dev())
    assert voting.dbscan_fit_predict(z).tolist() == [0] * 7                  # coincident points: one cluster
    assert voting.dbscan_fit_predict(z[:4]).tolist() == [-1] * 4             # below min_samples: noise
    assert voting.dbscan_fit_predict(z[:1]).tolist() == [-1]
    # one dense blob larger than several CTAs: every point neighbours hundreds of others (union-find contention)
    rng = np.random.default_rng(9)
    x = (rng.uniform(-1, 1, (3000, 3)) * np.array([1.0, 0.5, 0.3])).astype(np.float32)
    assert np.array_equal(voting.dbscan_fit_predict(t(x)).cpu().numpy(), O.dbscan(x))
    # row stride 4, foreground interleaved with the rest, sizes that are not multiples of the CTA
    n = 5003
    pts = np.concatenate([_blobs(rng, 6, 200), rng.uniform(-20, 20, (n, 3)).astype(np.float32)])[:n]
    pts = np.concatenate([pts, rng.uniform(0, 1, (n, 1)).astype(np.float32)], 1)[rng.permutation(n)]
    bf = rng.integers(0, 4, n).astype(np.int32)
    st = ops.cluster_boxes(t(pts), t(bf))
    fg = np.where(bf == 2)[0]
    m, n_clusters, n_kept = st["counts"].cpu().tolist()
    want = O.dbscan(pts[fg])
    kept, lo, hi = O.cluster_boxes(pts[fg], want)
    assert m == len(fg) and n_clusters == want.max() + 1 and n_kept == len(kept)
    assert np.array_equal(st["fg_index"][:m].cpu().numpy(), fg)
    assert np.array_equal(st["fg_label"][:m].cpu().numpy(), want)
    assert np.array_equal(st["kept_label"][:n_kept].cpu().numpy(), kept)
    assert np.array_equal(st["box_lo"][:n_kept].cpu().numpy(), lo)
    assert np.array_equal(st["box_hi"][:n_kept].cpu().numpy(), hi)


def test_cluster_vs_oracle_and_no_foreground():
    from streammos_b200 import voting
    rng = np.random.default_rng(21)
    objs = [rng.uniform(-30, 30, 3) * np.array([1, 1, 0.03]) + rng.uniform(-1, 1, (int(rng.integers(20, 300)), 3)) *
            np.array([0.8, 0.4, 0.3]) for _ in range(25)]
    objs.append(np.array([3.0, 3.0, -1.0]) + rng.uniform(-1, 1, (120, 3)) * np.array([0.8, 0.5, 0.04]))   # thin
    objs.append(np.array([-9.0, 4.0, -1.5]) + rng.uniform(-1, 1, (80, 3)) * np.array([0.5, 0.5, 0.0]))    # flat
    fgp = np.concatenate(objs)
    bg = rng.uniform(-40, 40, (20000, 3)) * np.array([1, 1, 0.05])
    cur = np.concatenate([fgp, bg]).astype(np.float32)
    bf = np.concatenate([np.full(len(fgp), 2), rng.integers(0, 2, len(bg))]).astype(np.int64)
    pred = rng.integers(0, 3, len(cur)).astype(np.int64)
    perm = rng.permutation(len(cur))
    cur, bf, pred = np.concatenate([cur, np.zeros((len(cur), 1), np.float32)], 1)[perm], bf[perm], pred[perm]
    local = np.concatenate([fgp + rng.normal(0, 0.05, fgp.shape) for _ in range(4)] +
                           [rng.uniform(-40, 40, (50000, 3)) * np.array([1, 1, 0.05])]).astype(np.float32)
    local = np.concatenate([local, np.zeros((len(local), 1), np.float32)], 1)
    lpred = rng.integers(0, 3, len(local)).astype(np.int64)
    want = O.cluster(cur, pred, bf, local, lpred)
    assert (want != pred).sum() > 500 and {1, 2} <= set(want[bf == 2].tolist())
    got = voting.cluster(t(cur), t(pred), t(bf), t(local), t(lpred))
    assert np.array_equal(got.cpu().numpy(), want)
    # no moving point: returned unchanged (voxel_instance_voting.py:146-147)
    same = voting.cluster(t(cur), t(pred), t(np.zeros_like(bf)), t(local), t(lpred))
    assert np.array_equal(same.cpu().numpy(), pred)
    with pytest.raises(RuntimeError):
        voting.cluster(t(cur), t(pred).to(torch.int32), t(bf), t(local), t(lpred))


def test_instance_vote_many_boxes_vs_oracle():
    from streammos_b200 import ops
    rng = np.random.default_rng(9)
    P, K = 300000, 300  # more boxes than one shared-memory chunk
    pts = np.concatenate([rng.uniform(-50, 50, (P, 2)), rng.uniform(-4, 2, (P, 1)), rng.uniform(0, 1, (P, 1))],
                         1).astype(np.float32)
    pred = rng.integers(0, 3, P).astype(np.int64)
    c = np.concatenate([rng.uniform(-45, 45, (K, 2)), rng.uniform(-3, 1, (K, 1))], 1)
    half = rng.uniform(0.5, 4, (K, 3))
    lo, hi = (c - half).astype(np.float32), (c + half).astype(np.float32)
    sums = ops.instance_vote(t(pts), t(pred), t(lo), t(hi))
    want = O.instance_vote(pts, pred, lo, hi)
    assert np.array_equal(sums.cpu().numpy(), want)
    # number of boxes read on the device: all of them (two chunks in one grid row), some, none
    for k in (K, 257, 40, 0):
        count = torch.tensor([k], dtype=torch.int32, device=dev())
        got = ops.instance_vote(t(pts), t(pred), t(lo), t(hi), count=count).cpu().numpy()
        assert np.array_equal(got[:k], want[:k]) and not got[k:].any(), k


# ------------------------------------------------------------------------------------------------
# Streaming long-term memory: pose alignment + crop + quantise + vote on resident scans
# ------------------------------------------------------------------------------------------------
def _fov_thresholds(size):
    fov, eps = ((-50, -50, -4), (50, 50, 2)), 1e-4
    lo = [np.float32(fov[0][i] + eps) for i in range(3)]
    hi = [np.float32(fov[1][i] - eps) for i in range(3)]
    mins = [float(fov[0][i]) for i in range(3)]
    deltas = [np.float32((fov[1][i] - fov[0][i]) / size[i]) for i in range(3)]
    return lo, hi, mins, deltas


def test_stream_vote_golden(golden):
    """The reference's own loop body (Trans, Crop, Quantize, voting, write-back) on a 9-scan synthetic drive."""
    from streammos_b200 import ops, voting
    g = golden("stream_vote_a")
    size = tuple(int(s) for s in g["size"])
    lo, hi, mins, deltas = _fov_thresholds(size)
    scans = [(t(g["scans"][j]), t(g["preds"][j]), g["pose_diffs"][j]) for j in range(8)]
    scans.append((t(g["scans"][8]), t(g["preds"][8]), None))
    vl, pl = ops.vote_stream(scans, 8, lo, hi, mins, deltas, size, 3)
    assert np.array_equal(vl.cpu().numpy(), g["voxel_labels"])
    assert np.array_equal(pl.cpu().numpy(), g["point_labels"])
    # the same frame through the ring-buffer object, from absolute poses (inv(pose_cur) . pose_hist on the host)
    sv = voting.StreamingVoter(frames_num_max=8, size=size)
    for j in list(range(7, -1, -1)) + [8]:                 # oldest first; fixture stores history newest first
        sv.push(t(g["scans"][j]), t(g["preds"][j]), g["poses"][j])
    vl2, pl2 = sv.vote()
    assert np.array_equal(vl2.cpu().numpy(), g["voxel_labels"])
    assert np.array_equal(pl2.cpu().numpy(), g["point_labels"])


def _drive(rng, n_scans, n):
    """Synthetic drive: ring-shaped scans, a turning ego trajectory, random predictions."""
    scans, preds, poses = [], [], []
    for k in range(n_scans):
        r = np.abs(rng.standard_normal(n)) * 20.0
        th = rng.uniform(0, 2 * np.pi, n)
        pts = np.stack([r * np.cos(th), r * np.sin(th), rng.normal(-1.5, 0.8, n), rng.uniform(0, 1, n)],
                       -1).astype(np.float32)
        yaw = 0.03 * k
        c, s_ = np.cos(yaw), np.sin(yaw)
        poses.append(np.array([[c, -s_, 0.001 * k, 1.1 * k], [s_, c, -0.002 * k, 0.02 * k * k],
                               [-0.001 * k, 0.002 * k, 1.0, 0.03 * k], [0, 0, 0, 1]], np.float64))
        scans.append(pts)
        preds.append(rng.integers(0, 3, n).astype(np.uint8))
    return scans, preds, poses


def test_stream_vote_sequence_config_size_vs_oracle():
    """A 12-scan stream at the config size (120k points, 512x512x30): the warm-up branch for the first 8 scans
    (each votes against the other 7, voxel_voting.py:195-214) and the sliding window afterwards (:177-194)."""
    from streammos_b200 import voting
    rng = np.random.default_rng(2024)
    n, size, hist = 120000, (512, 512, 30), 8
    lo, hi, mins, deltas = _fov_thresholds(size)
    scans, preds, poses = _drive(rng, 12, n)
    sv = voting.StreamingVoter(frames_num_max=hist, size=size)

    def oracle_frame(ids, cur):
        inv = np.linalg.inv(poses[cur])
        sc = [(scans[j], preds[j], inv.dot(poses[j]) if j != cur else None) for j in ids]
        return O.vote_stream(sc, ids.index(cur), lo, hi, mins, deltas, size, 3)

    for k in range(hist):
        sv.push(t(scans[k]), t(preds[k]), poses[k])
    changed = 0
    for k in (0, 3, 7):                                       # warm-up branch
        vl, pl = sv.vote(current=k)
        vl_ref, pl_ref = oracle_frame(list(range(hist)), k)
        assert np.array_equal(vl.cpu().numpy(), vl_ref) and np.array_equal(pl.cpu().numpy(), pl_ref)
    for k in range(hist, 12):                                 # steady state
        sv.push(t(scans[k]), t(preds[k]), poses[k])
        vl, pl = sv.vote()
        vl_ref, pl_ref = oracle_frame(list(range(k - hist, k + 1)), k)
        assert np.array_equal(vl.cpu().numpy(), vl_ref) and np.array_equal(pl.cpu().numpy(), pl_ref)
        changed += int((pl_ref != preds[k]).sum())
        outside = ~((scans[k][:, :3] > np.array(lo)) & (scans[k][:, :3] < np.array(hi))).all(1)
        assert outside.any() and np.array_equal(pl_ref[outside], preds[k][outside].astype(np.int64))
    assert changed > 0 and len(sv.ring) == hist + 1


def test_stream_vote_edge_cases():
    from streammos_b200 import ops
    size = (8, 8, 4)
    lo, hi, mins, deltas = _fov_thresholds(size)
    ident = np.eye(4)
    shift = np.eye(4); shift[0, 3] = 200.0                    # history pushed out of the crop box entirely
    cur = np.array([[0.5, 0.5, 0.0, 0], [49.99995, 0, 0, 0], [60, 0, 0, 0], [-49.9998, -49.9998, -3.9998, 0]], np.float32)
    cur_l = np.array([1, 2, 2, 2], np.uint8)
    h = np.array([[0.6, 0.6, 0.1, 0], [0.7, 0.4, 0.2, 0], [0.55, 0.45, 0.3, 0]], np.float32)
    h_l = np.array([2, 2, 0], np.uint8)
    empty = (torch.empty(0, 4, device=dev()), torch.empty(0, dtype=torch.uint8, device=dev()), ident)
    for scans_np in ([(h, h_l, ident), (cur, cur_l, None)], [(h, h_l, shift), (cur, cur_l, None)]):
        scans = [(t(a), t(b), m) for a, b, m in scans_np]
        for with_empty in (False, True):
            sc = ([empty] + scans) if with_empty else scans
            ci = len(sc) - 1
            vl, pl = ops.vote_stream(sc, ci, lo, hi, mins, deltas, size, 3)
            vl_ref, pl_ref = O.vote_stream(scans_np, 1, lo, hi, mins, deltas, size, 3)
            assert np.array_equal(vl.cpu().numpy(), vl_ref) and np.array_equal(pl.cpu().numpy(), pl_ref)
    # first case: voxel of point 0 holds labels {1, 2, 2, 0} -> 2; points on/outside the open box keep their label
    _, pl = ops.vote_stream([(t(h), t(h_l), ident), (t(cur), t(cur_l), None)], 1, lo, hi, mins, deltas, size, 3)
    assert pl.tolist()[0] == 2 and pl.tolist()[2] == 2
    with pytest.raises(Exception):
        ops.vote_stream([(t(cur).cpu(), t(cur_l), None)], 0, lo, hi, mins, deltas, size, 3)


# ------------------------------------------------------------------------------------------------
# PointNet stem (next: SURVEY §8f rank 4)
# ------------------------------------------------------------------------------------------------
def _stem_params(g):
    bn = [O.bn_affine(g["bn%d_weight" % i], g["bn%d_bias" % i], g["bn%d_mean" % i], g["bn%d_var" % i],
                      float(g["bn%d_eps" % i])) for i in range(3)]
    return bn[0], g["w1"], bn[1], g["w2"], bn[2]


@pytest.mark.parametrize("path", ["umma", "fma"])
def test_point_stem_golden(golden, path, monkeypatch):
    """Default: layer 2 on the tcgen05 tensor cores (3xTF32 split, fp32 accumulation in TMEM), inside the 1e-5 bar.
    SMOS_STEM_UMMA=0: the CUDA-core kernel, same FMA order as the oracle and bit-exact against it."""
    from streammos_b200 import ops
    from streammos_b200.backbone import PointNetStacker
    monkeypatch.setenv("SMOS_STEM_UMMA", "1" if path == "umma" else "0")
    g = golden("point_stem_a")
    bn0, w1, bn1, w2, bn2 = _stem_params(g)
    tt = lambda pair: (t(pair[0]), t(pair[1]))
    y = ops.point_stem_forward(t(g["x"]), tt(bn0), t(w1), tt(bn1), t(w2), tt(bn2))
    assert y.shape == g["out"].shape and y.is_contiguous()
    body, pads = slice(0, -50), slice(-50, None)
    want = O.point_stem(g["x"], bn0, w1, bn1, w2, bn2)
    if path == "fma":  # against the reference within fp32 rounding, against the oracle bit for bit
        assert np.array_equal(y[..., 0].cpu().numpy(), want)
    else:
        np.testing.assert_allclose(y[:, :, body, 0].cpu().numpy(), want[:, :, body], rtol=1e-5, atol=1e-5)
        ypm = ops.point_stem_forward(t(g["x"]), tt(bn0), t(w1), tt(bn1), t(w2), tt(bn2), point_major_out=True)
        assert ypm.stride(1) == 1 and torch.equal(ypm, y)      # point-major rows: same values, channels_last strides
    np.testing.assert_allclose(y[:, :, body, 0].cpu().numpy(), g["out"][:, :, body, 0], rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(y[:, :, pads, 0].cpu().numpy(), g["out64"][:, :, pads, 0], rtol=1e-5, atol=1e-3)
    # the drop-in module: the reference's state_dict loads, eval forward runs the fused kernel, train forward torch
    m = PointNetStacker(7, 64, pre_bn=True, stack_num=2)
    assert sorted(m.state_dict().keys()) == list(g["state_keys"])
    sd = {"layer.0.layer.1.weight": g["w1"], "layer.1.layer.0.weight": g["w2"]}
    for i, pre in enumerate(("layer.0.layer.0", "layer.0.layer.2", "layer.1.layer.1")):
        sd.update({pre + ".weight": g["bn%d_weight" % i], pre + ".bias": g["bn%d_bias" % i],
                   pre + ".running_mean": g["bn%d_mean" % i], pre + ".running_var": g["bn%d_var" % i],
                   pre + ".num_batches_tracked": np.int64(0)})
    m.load_state_dict({k: torch.as_tensor(v) for k, v in sd.items()})
    m = m.to(dev()).eval()
    from streammos_b200 import ops as _ops
    _ops.reset_launch_count()
    with torch.no_grad():
        ym = m(t(g["x"]))
    assert _ops.launch_count() == 1
    np.testing.assert_allclose(ym[:, :, body, 0].cpu().numpy(), g["out"][:, :, body, 0], rtol=1e-5, atol=1e-5)
    with torch.enable_grad():                                   # autograd / training go through the torch layers
        yt = m(t(g["x"]).requires_grad_(True))
    assert yt.requires_grad
    # torch runs these convolutions in TF32 on the GPU (cudnn.allow_tf32): only a loose check that the path works
    np.testing.assert_allclose(yt[:, :, body, 0].detach().cpu().numpy(), g["out"][:, :, body, 0], rtol=5e-2, atol=5e-2)


@pytest.mark.parametrize("path", ["umma", "fma"])
@pytest.mark.parametrize("B,Cin,N", [(3, 7, 120000), (1, 7, 1), (2, 5, 131), (1, 16, 4097)])
def test_point_stem_sizes_vs_oracle(B, Cin, N, path, monkeypatch):
    from streammos_b200 import ops
    monkeypatch.setenv("SMOS_STEM_UMMA", "1" if path == "umma" else "0")

    def same(got, want):
        if path == "fma":
            assert np.array_equal(got, want)
        else:
            np.testing.assert_allclose(got, want, rtol=1e-5, atol=2e-5)
    rng = np.random.default_rng(B * 1000 + N)
    x = rng.standard_normal((B, Cin, N, 1)).astype(np.float32) * 3
    w1 = (rng.standard_normal((64, Cin)) / np.sqrt(Cin)).astype(np.float32)
    w2 = rng.standard_normal((64, 64)).astype(np.float32) / 8
    bn = [(rng.uniform(0.5, 1.5, c).astype(np.float32), rng.standard_normal(c).astype(np.float32) * 0.2) for c in (Cin, 64, 64)]
    tt = lambda pair: (t(pair[0]), t(pair[1]))
    for bn0 in (bn[0], None):
        y = ops.point_stem_forward(t(x), tt(bn0) if bn0 else None, t(w1), tt(bn[1]), t(w2), tt(bn[2]))
        same(y[..., 0].cpu().numpy(), O.point_stem(x, bn0, w1, bn[1], w2, bn[2]))
    # strided input view (channels 0..Cin-1 of a wider tensor)
    wide = rng.standard_normal((B, Cin + 2, N, 1)).astype(np.float32)
    y = ops.point_stem_forward(t(wide)[:, :Cin], tt(bn[0]), t(w1), tt(bn[1]), t(w2), tt(bn[2]))
    same(y[..., 0].cpu().numpy(), O.point_stem(wide[:, :Cin], bn[0], w1, bn[1], w2, bn[2]))
    with pytest.raises(RuntimeError):
        ops.point_stem_forward(t(x).cpu(), None, t(w1), tt(bn[1]), t(w2), tt(bn[2]))
    with pytest.raises(NotImplementedError):                    # float64 weights are refused, not reinterpreted
        ops.point_stem_forward(t(x), None, t(w1).double(), tt(bn[1]), t(w2), tt(bn[2]))


# ------------------------------------------------------------------------------------------------
# Model input tensors from raw scans (next: SURVEY §8f rank 2, exact part)
# ------------------------------------------------------------------------------------------------
def test_form_batch_golden_and_config_size(golden):
    from streammos_b200 import ops, synthetic
    g = golden("form_batch_a")
    rng_ = ((-50.0, 50.0), (-50.0, 50.0), (-4.0, 2.0))
    for tag, (xs, ys) in {"pp": (1, 1), "mp": (-1, 1)}.items():
        feat, coord = ops.form_batch(t(g["points"]), *rng_, (512, 512, 30), xs, ys)
        assert feat.shape == (3, 7, 3000, 1) and coord.shape == (3, 3000, 3, 1)
        assert np.array_equal(feat[..., 0].cpu().numpy(), g["feat_" + tag])      # bit-exact with the reference's numpy
        assert np.array_equal(coord[..., 0].cpu().numpy(), g["coord_" + tag])
    s = synthetic.make_scan(77, 120000, 3)                                      # config size, vs the oracle
    feat, coord = ops.form_batch(t(s["xyzi"]), *rng_, (512, 512, 30))
    rf, rc = O.form_batch(s["xyzi"], *rng_, (512, 512, 30))
    assert np.array_equal(feat[..., 0].cpu().numpy(), rf) and np.array_equal(coord[..., 0].cpu().numpy(), rc)
    assert np.array_equal(rc[..., None], s["pcds_coord"])                       # = the harness's own loader restatement
    with pytest.raises(RuntimeError):
        ops.form_batch(t(s["xyzi"]).cpu(), *rng_, (512, 512, 30))


def test_quantize_torch_cuda_arithmetic_matches_torch_on_the_device():
    """the reference's voting scripts evaluate Quantize on CUDA tensors (voxel_voting.py:218-240),
    where torch turns `tensor / python_scalar` into a multiplication by the float32 reciprocal. arithmetic="torch_cuda"
    reproduces THAT bit for bit (checked against torch itself on this device, with the reference's three lines,
    voxel_voting.py:86-88); the default "ieee" mode is numpy's / torch-CPU's division and equals the oracle."""
    from streammos_b200 import synthetic, voting
    rx, ry, rz, size = (-50.0, 50.0), (-50.0, 50.0), (-4.0, 2.0), (512, 512, 30)
    pts = np.concatenate([synthetic.make_scan(70 + i, 120000, 1)["xyzi"][0] for i in range(9)])     # the 9-scan local map
    pcds = t(pts)
    dx, dy, dz = (rx[1] - rx[0]) / size[0], (ry[1] - ry[0]) / size[1], (rz[1] - rz[0]) / size[2]
    ref = torch.stack(((pcds[:, 0] - rx[0]) / dx, (pcds[:, 1] - ry[0]) / dy, (pcds[:, 2] - rz[0]) / dz), dim=-1)
    got = voting.Quantize(pcds, rx, ry, rz, size, arithmetic="torch_cuda")
    assert torch.equal(got, ref)
    assert torch.equal(got.to(torch.int64), ref.to(torch.int64))
    ieee = voting.Quantize(pcds, rx, ry, rz, size)
    assert np.array_equal(ieee.cpu().numpy(), O.quantize(pts, rx, ry, rz, size))
    cpu = torch.from_numpy(pts)
    ref_cpu = torch.stack(((cpu[:, 0] - rx[0]) / dx, (cpu[:, 1] - ry[0]) / dy, (cpu[:, 2] - rz[0]) / dz), dim=-1)
    assert torch.equal(ieee.cpu(), ref_cpu)                      # torch on the CPU divides
    diff = (ieee != got).any(1)
    assert 0 < int(diff.sum())                                   # the two arithmetics are not the same function ...
    assert bool(((ieee - got).abs() <= 2.0 ** -22 * ieee.abs().clamp(min=1e-30)).all())   # ... but at most an ulp or two apart
    with pytest.raises(ValueError):
        voting.Quantize(pcds, rx, ry, rz, size, arithmetic="fast")


def test_sphere_quantize_golden_and_config_size(golden):
    """utils.SphereQuantize on the device (smos_sphere_quantize). Floating point: the angles are float64 arctan2 / arcsin
    rounded to float32, so the kernel and the oracle agree to the last bit (up to a double-rounding case in ~1e-8 of the
    points) and both sit within 1 ulp of the angle of the reference's numpy output — stated in range-view cells."""
    from streammos_b200 import ops, synthetic
    tol_theta, tol_phi = 5e-5, 5e-4
    g = golden("sphere_a")
    for tag, (xs, ys) in {"pp": (1, 1), "mp": (-1, 1), "pm": (1, -1)}.items():
        got = ops.sphere_quantize(t(g["points"][None]), x_sign=xs, y_sign=ys)
        assert got.shape == (1, 6000, 2, 1)
        got = got[0, :, :, 0].cpu().numpy()
        ref = g["sphere_" + tag]
        d = np.abs(got.astype(np.float64) - ref)
        assert d[:, 0].max() <= tol_theta and d[:, 1].max() <= tol_phi
        assert np.array_equal(got[:13], ref[:13]) and np.array_equal(got[5600:], ref[5600:])
        assert np.array_equal(got, O.sphere_quantize(g["points"], x_sign=xs, y_sign=ys))
        for s in (1.0, 0.5, 0.25):
            assert np.array_equal(np.trunc(got * np.float32(s)), np.trunc(ref * np.float32(s)))
    s = synthetic.make_scan(78, 120000, 3)                                      # config size, all T frames, vs the oracle
    got = ops.sphere_quantize(t(s["xyzi"]))[..., 0].cpu().numpy()
    ref = O.sphere_quantize(s["xyzi"])
    assert (got != ref).mean() < 1e-5 and np.abs(got.astype(np.float64) - ref).max() <= tol_phi
    # against numpy's own float32 sequence on this host (the loader): the stated bound, and almost no cell changes
    x, y, z = s["xyzi"][..., 0], s["xyzi"][..., 1], s["xyzi"][..., 2]
    c = ops.sphere_constants()
    dist = np.sqrt(x ** 2 + y ** 2 + z ** 2) + np.float32(1e-12)
    host = np.stack(((np.float32(c[1]) - np.arcsin(z / dist)) / np.float32(c[3]),
                     (np.float32(c[0]) - np.arctan2(x, y)) / np.float32(c[2])), -1)
    assert host.dtype == np.float32
    d = np.abs(got.astype(np.float64) - host)
    assert d[..., 0].max() <= tol_theta and d[..., 1].max() <= tol_phi
    assert (np.trunc(got) != np.trunc(host)).any(-1).mean() < 1e-4
    with pytest.raises(RuntimeError):
        ops.sphere_quantize(t(s["xyzi"]).cpu())


def test_point_stem_tensor_core_variant(golden, monkeypatch):
    """SMOS_STEM_TC=1: layer 2 as 3xTF32 split products on mma.sync. Not bit-identical to the scalar FMA order, but
    inside the same 1e-5 bar against the reference module and the oracle (fixture + config size, raw and loader input)."""
    from streammos_b200 import ops, synthetic
    g = golden("point_stem_a")
    bn0, w1, bn1, w2, bn2 = _stem_params(g)
    tt = lambda pair: (t(pair[0]), t(pair[1]))
    monkeypatch.setenv("SMOS_STEM_UMMA", "0")
    scalar = ops.point_stem_forward(t(g["x"]), tt(bn0), t(w1), tt(bn1), t(w2), tt(bn2))
    monkeypatch.setenv("SMOS_STEM_TC", "1")
    y = ops.point_stem_forward(t(g["x"]), tt(bn0), t(w1), tt(bn1), t(w2), tt(bn2))
    body, pads = slice(0, -50), slice(-50, None)
    np.testing.assert_allclose(y[:, :, body, 0].cpu().numpy(), g["out"][:, :, body, 0], rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(y[:, :, pads, 0].cpu().numpy(), g["out64"][:, :, pads, 0], rtol=1e-5, atol=1e-3)
    np.testing.assert_allclose(y[:, :, body].cpu().numpy(), scalar[:, :, body].cpu().numpy(), rtol=1e-5, atol=1e-5)
    rng_ = ((-50.0, 50.0), (-50.0, 50.0), (-4.0, 2.0))
    pts = synthetic.make_scan(6, 120000, 3)["xyzi"]
    got, gc = ops.point_stem_forward_raw(t(pts), *rng_, (512, 512, 30), tt(bn0), t(w1), tt(bn1), t(w2), tt(bn2))
    rf, rc = O.form_batch(pts, *rng_, (512, 512, 30))
    ref = O.point_stem(rf, bn0, w1, bn1, w2, bn2)
    valid = np.abs(pts[..., 0]) < 100                                           # loader pads: see the golden test
    gn = got[..., 0].cpu().numpy()
    np.testing.assert_allclose(gn.transpose(0, 2, 1)[valid], ref.transpose(0, 2, 1)[valid], rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(gn, ref, rtol=1e-5, atol=2e-3)
    assert np.array_equal(gc[..., 0].cpu().numpy(), rc)


@pytest.mark.parametrize("path", ["umma", "fma"])
def test_point_stem_from_raw_points_equals_form_batch_then_stem(golden, path, monkeypatch):
    """The fused raw-point stem (smos_point_stem_forward_raw) is bit-identical to form_batch followed by the stem."""
    from streammos_b200 import ops, synthetic
    monkeypatch.setenv("SMOS_STEM_UMMA", "1" if path == "umma" else "0")
    g = golden("point_stem_a")
    bn = [O.bn_affine(g["bn%d_weight" % i], g["bn%d_bias" % i], g["bn%d_mean" % i], g["bn%d_var" % i],
                      float(g["bn%d_eps" % i])) for i in range(3)]
    tt = lambda pair: (t(pair[0]), t(pair[1]))
    rng_ = ((-50.0, 50.0), (-50.0, 50.0), (-4.0, 2.0))
    for n, (xs, ys) in ((120000, (1, 1)), (4099 * 4, (-1, 1)), (1, (1, -1))):
        pts = synthetic.make_scan(5, max(n, 64), 3)["xyzi"][:, :n]
        f7, c = ops.form_batch(t(pts), *rng_, (512, 512, 30), xs, ys)
        want = ops.point_stem_forward(f7, tt(bn[0]), t(g["w1"]), tt(bn[1]), t(g["w2"]), tt(bn[2]))
        got, gc = ops.point_stem_forward_raw(t(pts), *rng_, (512, 512, 30), tt(bn[0]), t(g["w1"]), tt(bn[1]), t(g["w2"]),
                                             tt(bn[2]), xs, ys)
        assert torch.equal(got, want) and torch.equal(gc, c)
    rf, rc = O.form_batch(pts, *rng_, (512, 512, 30), xs, ys)
    if path == "fma":
        assert np.array_equal(got[..., 0].cpu().numpy(), O.point_stem(rf, bn[0], g["w1"], bn[1], g["w2"], bn[2]))
    else:
        np.testing.assert_allclose(got[..., 0].cpu().numpy(), O.point_stem(rf, bn[0], g["w1"], bn[1], g["w2"], bn[2]),
                                   rtol=1e-5, atol=2e-3)


def test_point_stem_cta_cap_is_a_scheduling_hint_only(golden):
    """smos_point_stem_forward_raw_capped: any cap on the persistent CTAs gives the bits of the uncapped launch."""
    from streammos_b200 import ops, synthetic
    g = golden("point_stem_a")
    bn0, w1, bn1, w2, bn2 = _stem_params(g)
    tt = lambda pair: (t(pair[0]), t(pair[1]))
    rng_ = ((-50.0, 50.0), (-50.0, 50.0), (-4.0, 2.0))
    pts = t(synthetic.make_scan(9, 120000, 3)["xyzi"])
    args = (pts, *rng_, (512, 512, 30), tt(bn0), t(w1), tt(bn1), t(w2), tt(bn2))
    y0, c0 = ops.point_stem_forward_raw(*args, point_major_out=True)
    for cap in (1, 37, 90, 147, 100000):
        y, c = ops.point_stem_forward_raw(*args, point_major_out=True, max_ctas=cap)
        assert torch.equal(y, y0) and torch.equal(c, c0), cap


def test_step_from_raw_scan_matches_step_from_loader_tensors():
    """RawBatch (Quantize + make_point_feat on the device) and LoaderBatch (done by the host) give the same step."""
    from streammos_b200 import stream
    n = 20000
    a, b = stream.HotPath(dev(), n_points=n, seed=4), stream.HotPath(dev(), n_points=n, seed=4)
    with torch.no_grad():
        for i in range(2):
            la, sa, pa = a.step(stream.make_host_loader_scan(800 + i, n, pin=False).to(dev()))
            lb, sb, pb = b.step(stream.make_host_raw_scan(800 + i, n, pin=False).to(dev()))
            assert torch.equal(la, lb) and torch.equal(sa, sb)
            for x, y in zip(pa, pb):
                assert torch.equal(x, y)


# ------------------------------------------------------------------------------------------------
# Whole hot path: streaming harness on the GPU vs the CPU restatement of the same sequence
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("point_major", [True, False])
def test_whole_path_stream_matches_cpu_path(point_major):
    """config #2/#3/#1 chained as one scan step (5 pools, 5 gathers, 2 deformable-attention layers, voxel +
    instance voting) for three consecutive scans with carried short- and long-term memory."""
    from oracle.cpu_path import CpuHotPath
    from streammos_b200 import stream
    n = 30000
    hot = stream.HotPath(dev(), n_points=n, seed=5, point_major=point_major)
    cpu = stream.HotPath("cpu", n_points=n, seed=5)
    state = {"x0": cpu.x0, "x1": cpu.x1, "dec": cpu.dec, "memory": cpu.memory.clone(),
             "local_pts": cpu.local_pts.clone(), "local_pred": cpu.local_pred.clone(), "box_lo": cpu.box_lo,
             "box_hi": cpu.box_hi, "scan_index": 0}
    ref = CpuHotPath(state)
    with torch.no_grad():
        for i in range(3):
            scan = stream.make_host_scan(100 + i, n, pin=False)
            labels, sums, proj = hot.step(scan.to(dev()))
            r_labels, r_sums, r_proj = ref.step(scan)
            assert torch.equal(labels.cpu(), r_labels)                       # voted point labels: bit-exact
            assert np.array_equal(sums.cpu().numpy(), r_sums)                # instance votes: bit-exact
            assert torch.equal(proj[0].cpu(), r_proj[0])                     # pool #1 (pure scatter-max): bit-exact
            for a, b in zip(proj[1:], r_proj[1:]):                           # chains through bilinear gathers
                torch.testing.assert_close(a.cpu(), b, rtol=1e-4, atol=1e-5)
            torch.testing.assert_close(hot.memory.cpu(), state["memory"], rtol=1e-4, atol=1e-5)


def test_step_from_loader_tensors_matches_explicit_inputs():
    """A step that starts from the loader's tensors (7-channel point features, (T,N,3,1) coordinates used through a
    strided [:, :, :2] view, raw points taken from pcds_xyzi) equals the step on explicitly materialised inputs."""
    from streammos_b200 import stream
    n = 20000
    a = stream.HotPath(dev(), n_points=n, seed=3)
    b = stream.HotPath(dev(), n_points=n, seed=3)
    with torch.no_grad():
        for i in range(2):
            lb = stream.make_host_loader_scan(700 + i, n, pin=False).to(dev())
            sb = stream.ScanBatch(feat=a.point_pre(lb.pcds_xyzi), coord_bev=lb.pcds_coord[:, :, :2].contiguous(),
                                  coord_rv=lb.pcds_sphere_coord[:1].contiguous(),
                                  xyzi=lb.pcds_xyzi[0, :4, :, 0].t().contiguous(), pred=lb.pred, loc=lb.loc, attn=lb.attn)
            la, sa, pa = a.step(sb)
            l2, s2, p2 = b.step(lb)
            assert torch.equal(la, l2) and torch.equal(sa, s2)
            for x, y in zip(pa, p2):
                assert torch.equal(x, y)
            assert not lb.coord_bev.is_contiguous() and float(pa[0].abs().sum()) > 0


@pytest.mark.parametrize("use_graphs", [True, False])
def test_pipeline_matches_serial_step(use_graphs):
    """ScanPipeline (3 graphs per scan, default scans in flight) == the serial step, scan by scan."""
    from streammos_b200 import pipeline, stream
    n, n_buf, n_scans = 20000, 8, 12  # graph replays bake the voting ring slot: buffers = a multiple of 8
    host = [stream.make_host_scan(300 + j, n, pin=False) for j in range(n_buf)]
    serial = stream.HotPath(dev(), n_points=n, seed=9)
    want = []
    with torch.no_grad():
        for i in range(n_scans):
            labels, sums, _ = serial.step(host[i % n_buf].to(dev()))
            want.append((labels.clone(), sums.clone()))
    torch.cuda.synchronize()
    hot = stream.HotPath(dev(), n_points=n, seed=9)
    state0 = (hot.memory.clone(), hot.local_pts.clone(), hot.local_pred.clone())
    devb = [h.to(dev()) for h in host]
    pipe = pipeline.ScanPipeline(hot, devb, use_graphs=use_graphs)
    # the constructor's warm-up advanced the memories: restore the initial state before the real run
    hot.memory.copy_(state0[0]); hot.local_pts.copy_(state0[1]); hot.local_pred.copy_(state0[2])
    hot.scan_index = 0
    torch.cuda.synchronize()
    got = []
    for i in range(n_scans):
        j = pipe.submit()
        pipe.join(torch.cuda.current_stream())
        torch.cuda.synchronize()
        got.append((pipe.out[j][0].clone(), pipe.out[j][1].clone()))
    for i, ((wl, ws), (gl, gs)) in enumerate(zip(want, got)):
        assert torch.equal(wl, gl), "labels differ at scan %d" % i
        assert torch.equal(ws, gs), "instance votes differ at scan %d" % i
    torch.testing.assert_close(hot.memory, serial.memory, rtol=0, atol=0)


@pytest.mark.parametrize("in_flight", [2, 4, 8])
def test_pipeline_with_scans_really_in_flight(in_flight):
    """Eight scans submitted back to back (no host join in between: up to `in_flight` projections overlap on their own
    streams, temporal fusion and voting stay ordered by events) give the serial results, scan by scan."""
    from streammos_b200 import pipeline, stream
    n, n_buf = 20000, 8
    host = [stream.make_host_scan(700 + j, n, pin=False) for j in range(n_buf)]
    serial = stream.HotPath(dev(), n_points=n, seed=4)
    want = []
    with torch.no_grad():
        for i in range(n_buf):
            labels, sums, _ = serial.step(host[i].to(dev()))
            want.append((labels.clone(), sums.clone()))
    hot = stream.HotPath(dev(), n_points=n, seed=4)
    state0 = (hot.memory.clone(), hot.local_pts.clone(), hot.local_pred.clone())
    pipe = pipeline.ScanPipeline(hot, [h.to(dev()) for h in host], use_graphs=True, scans_in_flight=in_flight)
    hot.memory.copy_(state0[0]); hot.local_pts.copy_(state0[1]); hot.local_pred.copy_(state0[2])
    hot.scan_index = 0
    torch.cuda.synchronize()
    for i in range(n_buf):
        pipe.submit()
    pipe.join(torch.cuda.current_stream())
    torch.cuda.synchronize()
    for j, (wl, ws) in enumerate(want):
        assert torch.equal(wl, pipe.out[j][0]), "labels differ at scan %d" % j
        assert torch.equal(ws, pipe.out[j][1]), "instance votes differ at scan %d" % j
    torch.testing.assert_close(hot.memory, serial.memory, rtol=0, atol=0)


# ------------------------------------------------------------------------------------------------
# Scan ingestion: pose alignment + range filter + compaction + padding on the device
# ------------------------------------------------------------------------------------------------
def _ingest_inputs(g):
    return [(g["raw"][tt, :int(g["n_raw"][tt])], None if np.isnan(g["pose_diff"][tt]).any() else g["pose_diff"][tt])
            for tt in range(len(g["n_raw"]))]


def test_ingest_frames_golden(golden):
    """smos_ingest_frames against the loader's own utils.Trans / filter_pcds_mask / padding (tests/golden/ingest_a.npz),
    bit for bit; frame capacities larger than the point counts (the counts are read on the device)."""
    from streammos_b200 import ops
    g = golden("ingest_a")
    n_out = int(g["n_out"])
    frames = [(t(g["raw"][k]), int(g["n_raw"][k]), p) for k, (_, p) in enumerate(_ingest_inputs(g))]
    out, cnt, src = ops.ingest_frames(frames, (-50.0, 50.0), (-50.0, 50.0), (-4.0, 2.0), n_out, want_src=True)
    assert np.array_equal(cnt.cpu().numpy(), g["count"])
    assert np.array_equal(out.cpu().numpy().view(np.uint32), g["out"].view(np.uint32))
    want_out, want_cnt, want_src = O.ingest_frames(_ingest_inputs(g), (-50.0, 50.0), (-50.0, 50.0), (-4.0, 2.0), n_out)
    assert np.array_equal(src.cpu().numpy(), want_src)


@pytest.mark.parametrize("n_raw,n_out", [(131072, 120000), (1000, 4096), (5000, 1024), (1, 8), (4097, 4097)])
def test_ingest_frames_vs_oracle(n_raw, n_out):
    """Config size (131 k raw points -> 120 k rows), output longer / shorter than the input (a frame that does not fit
    is truncated and the count says so), single point, tile edges; device-resident counts and poses."""
    from streammos_b200 import ops
    rng = np.random.default_rng(n_raw + n_out)
    frames_np = []
    for k in range(3):
        n = max(1, n_raw - 37 * k)
        raw = np.stack([rng.uniform(-60, 60, n), rng.uniform(-60, 60, n), rng.uniform(-5, 3, n), rng.uniform(0, 1, n)],
                       -1).astype(np.float32)
        if n_out >= 100000:   # LiDAR-like: most points inside the range, so that the frame fits as in the loader
            raw[:, :2] *= np.float32(0.7)
            raw[:, 2] = rng.uniform(-3.9, 1.9, n).astype(np.float32)
        a = 0.01 * (k + 1)
        pose = np.eye(4)
        pose[:2, :2] = [[np.cos(a), -np.sin(a)], [np.sin(a), np.cos(a)]]
        pose[:3, 3] = (0.6 * k, -0.04 * k, 0.01 * k)
        frames_np.append((raw, pose if k else None))
    want_out, want_cnt, want_src = O.ingest_frames(frames_np, (-50.0, 50.0), (-50.0, 50.0), (-4.0, 2.0), n_out)
    cap = n_raw + 100
    frames = []
    for raw, pose in frames_np:
        buf = np.full((cap, 4), 7.0, np.float32)   # rows behind n hold in-range garbage that must be ignored
        buf[:len(raw)] = raw
        frames.append((t(buf), torch.tensor([len(raw)], dtype=torch.int32, device=dev()),
                       None if pose is None else t(np.ascontiguousarray(pose[:3]).reshape(12))))
    out, cnt, src = ops.ingest_frames(frames, (-50.0, 50.0), (-50.0, 50.0), (-4.0, 2.0), n_out, want_src=True)
    assert np.array_equal(cnt.cpu().numpy(), want_cnt)
    assert np.array_equal(out.cpu().numpy().view(np.uint32), want_out.view(np.uint32))
    assert np.array_equal(src.cpu().numpy(), want_src)


def test_resident_window_step_matches_host_aligned_frames():
    """The stream harness with the raw scans of the T-frame window resident in HBM (pose alignment + range filter +
    padding on the device) gives exactly the step the host-aligned frames give: same pooled grids, same labels."""
    from streammos_b200 import stream
    n, scans = 120000, 4
    resident, aligned = stream.make_host_resident_stream(3, scans, n, pin=False)
    dres = stream.link_window([b.to(dev()) for b in resident])
    dali = [b.to(dev()) for b in aligned]
    hot_a = stream.HotPath(dev(), n, seed=5)
    hot_b = stream.HotPath(dev(), n, seed=5)
    with torch.no_grad():
        for i in range(scans):
            la, sa, pa = hot_a.step(dres[i])
            lb, sb, pb = hot_b.step(dali[i])
            assert torch.equal(dres[i].points, dali[i].points)          # the ingested frames ARE the loader's frames
            assert int(dres[i].n_valid[0]) < n
            assert torch.equal(la, lb) and torch.equal(sa, sb)
            for x, y in zip(pa, pb):
                assert torch.equal(x, y)
    assert torch.equal(hot_a.memory, hot_b.memory)


def test_resident_window_with_device_sphere_quantize():
    """The same stream with NOTHING from the loader on the host side: the range-view coordinates of the current frame
    come from smos_sphere_quantize on the ingested frame. Floating point against numpy's float32 arctan2 / arcsin, so the
    bar is the stated one: coordinates within 1 ulp of the angle (in cells), at most 1e-4 of the points in another
    range-view cell; everything that does not read range-view coordinates (pool #1, voting, instance votes) identical."""
    from streammos_b200 import ops, stream, synthetic
    n, scans = 120000, 3
    resident, _ = stream.make_host_resident_stream(3, scans, n, pin=False)
    bare, _ = stream.make_host_resident_stream(3, scans, n, pin=False, device_sphere=True)
    assert bare[0].nbytes() == resident[0].nbytes() - n * 8 and bare[0].coord_rv is None
    dres = stream.link_window([b.to(dev()) for b in resident])
    dbar = stream.link_window([b.to(dev()) for b in bare])
    hot_a = stream.HotPath(dev(), n, seed=5)
    hot_b = stream.HotPath(dev(), n, seed=5, sphere_on_device=True)
    with torch.no_grad():
        for i in range(scans):
            la, sa, pa = hot_a.step(dres[i])
            lb, sb, pb = hot_b.step(dbar[i])
            assert torch.equal(dres[i].points, dbar[i].points)
            assert torch.equal(la, lb) and torch.equal(sa, sb) and torch.equal(pa[0], pb[0])
            got = ops.sphere_quantize(dbar[i].points[:1], theta_range=synthetic.RV_THETA, size=synthetic.RV_SHAPE)
            d = (got.double() - dres[i].coord_rv.double()).abs()[0, :, :, 0]
            assert float(d[:, 0].max()) <= 5e-5 and float(d[:, 1].max()) <= 5e-4
            moved = (torch.trunc(got) != torch.trunc(dres[i].coord_rv)).any(2).float().mean()
            assert float(moved) < 1e-4
            # downstream of the range view the bilinear gathers sample <= 5e-4 cells off: values move by that times the
            # local slope everywhere, and by more only where a point changed its cell
            for x, y in zip(pa[1:], pb[1:]):
                assert x.shape == y.shape and float(((x - y).abs() > 5e-3).float().mean()) < 2e-3
