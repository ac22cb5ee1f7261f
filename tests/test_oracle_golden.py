"""Pin the CPU oracle (oracle/smos_oracle.c) against outputs of the reference itself
(tests/golden/*.npz, made by tools/make_golden.py from /root/reference). Runs anywhere, no GPU."""
import numpy as np
import pytest

from oracle import oracle as O

POOL = ["pool_a", "pool_b", "pool_c"]
BILINEAR = ["bilinear_a", "bilinear_b"]
MSDA = ["msda_reftest", "msda_reftest_d30", "msda_reftest_d32", "msda_reftest_d71", "msda_streammos_small"]


@pytest.mark.parametrize("name", POOL)
def test_pool_forward_bit_exact(golden, name):
    g = golden(name)
    out, idx = O.voxel_maxpool_forward(g["feat"], g["ind"], (int(g["H"]), int(g["W"])), g["scale"], want_idx=True)
    assert np.array_equal(out, g["out"])  # max is exact: bit equality with point_deep.cpp
    # voxel_max_idx is consistent with the output: valid points never exceed their cell's value
    B, C, N = g["feat"].shape[:3]
    flat = out.reshape(-1)
    feat = g["feat"].reshape(B, C, N)
    valid = idx >= 0
    assert valid.any() and (~valid).any()
    for b in range(B):
        v = np.nonzero(valid[b])[0]
        for c in (0, C - 1):
            assert (flat[idx[b, v] + c * out.shape[2] * out.shape[3]] >= feat[b, c, v]).all()


@pytest.mark.parametrize("name", POOL)
def test_pool_backward_bit_exact(golden, name):
    g = golden(name)
    gf = O.voxel_maxpool_backward(g["feat"], g["ind"], g["out"], g["gout"], g["scale"])
    assert np.array_equal(gf, g["gfeat"].reshape(gf.shape))


@pytest.mark.parametrize("name", BILINEAR)
def test_bilinear_forward(golden, name):
    g = golden(name)
    out = O.bilinear_sample(g["grid"], g["coord"], g["scale"])
    ref = g["out"][..., 0]
    # north_star tolerance: 1e-5 relative in fp32 (+ small atol near zero crossings)
    np.testing.assert_allclose(out, ref, rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(out, g["out64"][..., 0], rtol=1e-4, atol=1e-5)


@pytest.mark.parametrize("name", BILINEAR)
def test_bilinear_backward(golden, name):
    g = golden(name)
    H, W = g["grid"].shape[2:]
    gg = O.bilinear_sample_backward(g["gout"], g["coord"], g["scale"], H, W)
    np.testing.assert_allclose(gg, g["ggrid"], rtol=1e-4, atol=1e-4)


@pytest.mark.parametrize("name", MSDA)
def test_msda_forward(golden, name):
    g = golden(name)
    out = O.ms_deform_attn_forward(g["value"], g["shapes"], g["lsi"], g["loc"], g["attn"])
    # deformattn/test.py:41 uses torch.allclose defaults in double
    np.testing.assert_allclose(out, g["out64"], rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(out, g["out32"], rtol=1e-2, atol=1e-3)  # test.py:56 float tolerance


@pytest.mark.parametrize("name", MSDA)
def test_msda_backward(golden, name):
    g = golden(name)
    gv, gl, ga = O.ms_deform_attn_backward(g["value"], g["shapes"], g["lsi"], g["loc"], g["attn"], g["gout"])
    np.testing.assert_allclose(gv, g["gvalue"], rtol=1e-8, atol=1e-11)
    np.testing.assert_allclose(ga, g["gattn"], rtol=1e-8, atol=1e-11)
    np.testing.assert_allclose(gl, g["gloc"], rtol=1e-7, atol=1e-10)


@pytest.mark.parametrize("name", ["msda_module_streammos", "msda_module_boxes"])
def test_msda_fused_front_end(golden, name):
    """Softmax + sampling-location arithmetic of the reference MODULE (ms_deform_attn.py:96-108), restated by
    oracle.ms_deform_attn_fused_forward, against the tensors the module itself handed to its sampling core."""
    g = golden(name)
    out = O.ms_deform_attn_fused_forward(g["value"], g["shapes"], g["lsi"], g["offsets"], g["logits"], g["ref"])
    np.testing.assert_allclose(out, g["core_out"], rtol=1e-4, atol=1e-5)   # fixture is the module's float32 run
    ref = O.ms_deform_attn_forward(g["value"], g["shapes"], g["lsi"], g["loc"], g["attn"])
    np.testing.assert_allclose(out, ref, rtol=1e-4, atol=1e-5)


def test_voting_bit_exact(golden):
    g = golden("voting_a")
    size = tuple(int(s) for s in g["size"])
    q = O.quantize(g["pts"], g["rx"], g["ry"], g["rz"], size)
    assert np.array_equal(q, g["quan"])
    coords = q.astype(np.int64)  # .to(torch.int64): truncation
    assert np.array_equal(coords, g["coords"])
    vl = O.determine_voxel_labels(coords, g["labels"], size)
    assert np.array_equal(vl, g["voxel_labels"])
    assert np.array_equal(O.determine_voxel_labels(coords, g["labels"], size, num_classes=3), vl)
    pl = O.get_point_labels_from_voxel_labels(g["cur"], vl, size)
    assert np.array_equal(pl, g["point_labels"])
    assert (vl > 0).any() and (pl == 0).any()


def test_instance_vote_bit_exact(golden):
    g = golden("instance_a")
    lo, hi = g["corners"].min(1), g["corners"].max(1)
    sums = O.instance_vote(g["local_pts"], g["local_pred"], lo, hi)
    assert np.array_equal(sums[:, 0], g["stat"])
    assert np.array_equal(sums[:, 1], g["dyn"])
    label = np.where(sums[:, 1] > sums[:, 0], 2, 1)
    assert np.array_equal(label, g["label"])
    assert set(label.tolist()) == {1, 2}  # the fixture exercises both outcomes


def test_cluster_matches_reference(golden):
    """DBSCAN labels as sklearn gave them inside the reference's cluster(), and cluster()'s relabelled output."""
    g = golden("cluster_a")
    fg = np.where(g["cur_bf"] == 2)[0]
    labels = O.dbscan(g["cur_pts"][fg])
    assert np.array_equal(labels, g["fg_labels"])
    sizes = np.bincount(labels[labels >= 0])
    assert (sizes <= 30).any() and (sizes == 31).any() and (labels == -1).any()   # both sides of the >30 cut
    out = O.cluster(g["cur_pts"], g["cur_pred"], g["cur_bf"], g["local_pts"], g["local_pred"])
    assert np.array_equal(out, g["cluster_out"])
    assert (out != g["cur_pred"]).sum() > 100
    # the older fixture carries a full cluster() run as well
    g = golden("instance_a")
    out = O.cluster(g["cur_pts"], g["cur_pred"], g["cur_bf"], g["local_pts"], g["local_pred"])
    assert np.array_equal(out, g["cluster_out"])
    # no moving point -> unchanged (voxel_instance_voting.py:146-147)
    none = O.cluster(g["cur_pts"], g["cur_pred"], np.zeros_like(g["cur_bf"]), g["local_pts"], g["local_pred"])
    assert np.array_equal(none, g["cur_pred"])


def test_dbscan_oracle_against_sklearn():
    """The third-party algorithm itself, where it is importable: random clouds with touching blobs, so that border
    points adjacent to two clusters and stolen-border clusters occur."""
    sk = pytest.importorskip("sklearn.cluster")
    rng = np.random.default_rng(5)
    for trial in range(6):
        blobs = [rng.uniform(-6, 6, 3) + rng.uniform(-1, 1, (int(rng.integers(5, 400)), 3)) * rng.uniform(0.2, 1.4, 3)
                 for _ in range(int(rng.integers(1, 7)))]
        blobs.append(rng.uniform(-8, 8, (int(rng.integers(0, 1500)), 3)))
        x = np.concatenate(blobs).astype(np.float32)
        x = x[rng.permutation(len(x))]
        want = sk.DBSCAN(eps=0.3, min_samples=5).fit_predict(x)
        assert np.array_equal(O.dbscan(x), want), trial
    # lattices whose axis distances sit right at eps (2 x 0.15f, 3 x 0.1f): the float64 reduced distance decides
    for step, depth, frac in ((0.15, 4, 0.55), (0.1, 3, 0.12)):
        g = np.arange(15, dtype=np.float32) * np.float32(step)
        x = np.stack(np.meshgrid(g, g, g[:depth], indexing="ij"), -1).reshape(-1, 3)
        x = x[rng.uniform(0, 1, len(x)) < frac]
        x = x[rng.permutation(len(x))]
        assert np.array_equal(O.dbscan(x), sk.DBSCAN(eps=0.3, min_samples=5).fit_predict(x)), step
    assert O.dbscan(np.zeros((0, 3), np.float32)).shape == (0,)
    assert np.array_equal(O.dbscan(np.zeros((7, 3), np.float32)), np.zeros(7, np.int32))   # coincident points
    assert np.array_equal(O.dbscan(np.zeros((4, 3), np.float32)), np.full(4, -1, np.int32))


def _stream_vote_setup(g):
    """(scans, crop_lo, crop_hi, mins, deltas, size) of the fixture, thresholds as StreamingVoter derives them."""
    size = tuple(int(s) for s in g["size"])
    fov, eps = ((-50, -50, -4), (50, 50, 2)), 1e-4
    lo = [np.float32(fov[0][i] + eps) for i in range(3)]
    hi = [np.float32(fov[1][i] - eps) for i in range(3)]
    mins = [float(fov[0][i]) for i in range(3)]
    deltas = [np.float32((fov[1][i] - fov[0][i]) / size[i]) for i in range(3)]
    n_hist = len(g["pose_diffs"])
    scans = [(g["scans"][j], g["preds"][j], g["pose_diffs"][j]) for j in range(n_hist)]
    scans.append((g["scans"][n_hist], g["preds"][n_hist], None))
    return scans, lo, hi, mins, deltas, size


def test_stream_vote_bit_exact(golden):
    """One frame of the voxel_voting.py loop (Trans + Crop + Quantize + vote + write-back), reference outputs."""
    g = golden("stream_vote_a")
    scans, lo, hi, mins, deltas, size = _stream_vote_setup(g)
    vl, pl, tout = O.vote_stream(scans, len(scans) - 1, lo, hi, mins, deltas, size, 3, want_transformed=True)
    n = g["scans"].shape[1]
    # float64 pose matmul stored as float32 (datasets/utils.py:116-126): bit-exact with numpy's dgemm
    assert np.array_equal(tout[: 8 * n].reshape(8, n, 3), g["transformed"][..., :3])
    assert np.array_equal(vl, g["voxel_labels"])
    assert np.array_equal(pl, g["point_labels"])
    assert 0 < int(g["n_cropped"]) < n                      # some current points lie outside the crop ...
    assert (pl != g["preds"][-1]).any()                     # ... and voting changes some labels inside it


def _stem_params(g):
    bn = [O.bn_affine(g["bn%d_weight" % i], g["bn%d_bias" % i], g["bn%d_mean" % i], g["bn%d_var" % i],
                      float(g["bn%d_eps" % i])) for i in range(3)]
    return bn[0], g["w1"], bn[1], g["w2"], bn[2]


def test_point_stem_matches_reference_module(golden):
    """PointNetStacker(7, 64, pre_bn=True, stack_num=2).eval() of the reference (networks/backbone.py:199-250)."""
    g = golden("point_stem_a")
    y = O.point_stem(g["x"], *_stem_params(g))
    ref, ref64 = g["out"][..., 0], g["out64"][..., 0]
    body, pads = slice(0, -50), slice(-50, None)
    # fp32, 1e-5 relative (+ 1e-5 absolute at the ReLU kink and for sums that cancel)
    np.testing.assert_allclose(y[..., body], ref[..., body], rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(y[..., body], ref64[..., body], rtol=1e-5, atol=1e-5)
    # loader pads (x = y = -1000, z = -4000): terms of 1e3..1e4 cancel, fp32 itself is only good to ~4e-4 there
    # (the reference's own fp32 forward differs from its fp64 forward by that much)
    assert np.abs(ref[..., pads] - ref64[..., pads]).max() > 1e-4
    np.testing.assert_allclose(y[..., pads], ref64[..., pads], rtol=1e-5, atol=1e-3)
    assert 0.2 < (ref > 0).mean() < 0.8   # both sides of the ReLU are exercised


def test_form_batch_bit_exact(golden):
    """Quantize + make_point_feat of the loader's form_batch (datasets/utils.py:151-169, data_StreamMOS.py:25-50,
    471-513), both TTA sign pairs of the fixture: float32 arithmetic replayed exactly."""
    g = golden("form_batch_a")
    for tag, (xs, ys) in {"pp": (1, 1), "mp": (-1, 1)}.items():
        feat, coord = O.form_batch(g["points"], (-50.0, 50.0), (-50.0, 50.0), (-4.0, 2.0), (512, 512, 30), xs, ys)
        assert np.array_equal(feat, g["feat_" + tag]) and np.array_equal(coord, g["coord_" + tag])
    assert float(g["feat_pp"][:, 4].min()) == np.float32(1e-12)          # points at the sensor origin
    assert (g["coord_pp"][..., 0] == 0).any() and (g["coord_pp"] < 0).any()  # lower bound hit, pads out of range


# SphereQuantize is a floating-point result (numpy's float32 arctan2 / arcsin are host specific in their last bit):
# the bar is 1 ulp of the ANGLE, expressed in range-view cells — the same numbers include/streammos_b200.h states
SPHERE_TOL_THETA, SPHERE_TOL_PHI = 5e-5, 5e-4


def test_sphere_quantize_within_one_ulp_of_the_angle(golden):
    """utils.SphereQuantize (datasets/utils.py:172-192) run by the reference itself vs the float64-angle restatement:
    coordinates within 1 ulp of the angle (in cells), the axis / origin / padding rows exact, and — on this fixture —
    no point changes its range-view cell at any of the three grids the model pools into."""
    g = golden("sphere_a")
    for tag, (xs, ys) in {"pp": (1, 1), "mp": (-1, 1), "pm": (1, -1)}.items():
        got = O.sphere_quantize(g["points"], x_sign=xs, y_sign=ys)
        ref = g["sphere_" + tag]
        assert got.shape == ref.shape == (6000, 2) and got.dtype == np.float32
        d = np.abs(got.astype(np.float64) - ref)
        assert d[:, 0].max() <= SPHERE_TOL_THETA and d[:, 1].max() <= SPHERE_TOL_PHI
        assert (d == 0).mean() > 0.75                          # most points are identical to the last bit
        assert np.array_equal(got[:13], ref[:13])              # origin and axis points: exact angles
        assert np.array_equal(got[5600:], ref[5600:])          # the loader's padding rows
        for s in (1.0, 0.5, 0.25):                             # cell rule of VoxelMaxPool: int(coord * scale)
            assert np.array_equal(np.trunc(got * np.float32(s)), np.trunc(ref * np.float32(s)))


def _ingest_frames(g):
    return [(g["raw"][t, :int(g["n_raw"][t])], None if np.isnan(g["pose_diff"][t]).any() else g["pose_diff"][t])
            for t in range(len(g["n_raw"]))]


def test_ingest_frames_bit_exact(golden):
    """Pose alignment + range filter + ordered compaction + padding (datasets/data_StreamMOS.py:515-574) against the
    loader's own utils.Trans / utils.filter_pcds_mask: every float32 of every frame, the counts and the masks."""
    g = golden("ingest_a")
    out, cnt, src = O.ingest_frames(_ingest_frames(g), (-50.0, 50.0), (-50.0, 50.0), (-4.0, 2.0), int(g["n_out"]))
    assert np.array_equal(cnt, g["count"])
    assert np.array_equal(out.view(np.uint32), g["out"].view(np.uint32))   # bit patterns, pads included
    for t in range(len(cnt)):
        assert np.array_equal(src[t, :cnt[t]], np.nonzero(g["mask"][t])[0]) and (src[t, cnt[t]:] == -1).all()
    # the frame without a pose hits the bounds exactly: lower bounds are inside, upper bounds are not
    keep = g["mask"][3][:8]
    assert keep.tolist() == [True, False, True, True, True, False, True, False]
