"""Host-side logic of the reference-shaped shims that needs no GPU: argument checks, error behaviour,
drop-in module registration."""
import sys

import pytest
import torch


def test_voxel_maxpool_asserts_like_the_reference():
    from streammos_b200 import deep_point
    feat = torch.zeros(1, 4, 10, 1)
    with pytest.raises(AssertionError):  # N mismatch, deep_point/__init__.py:21
        deep_point.VoxelMaxPool(feat, torch.zeros(1, 9, 2, 1), (4, 4), (1.0, 1.0))
    with pytest.raises(AssertionError):  # D != len(output_size), :22
        deep_point.VoxelMaxPool(feat, torch.zeros(1, 10, 2, 1), (4, 4, 4), (1.0, 1.0))
    with pytest.raises(AssertionError):  # dtype mismatch, :18
        deep_point.VoxelMaxPool(feat, torch.zeros(1, 10, 2, 1, dtype=torch.float64), (4, 4), (1.0, 1.0))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        deep_point.VoxelMaxPool(feat, torch.zeros(1, 10, 2, 1), (4, 4), (1.0, 1.0))


def test_cpu_kernel_stub_raises():
    from streammos_b200.point_deep import cpu_kernel
    with pytest.raises(RuntimeError):
        cpu_kernel.voxel_maxpooling_cpu_forward()


def test_msda_rejects_cpu_and_bad_im2col_step():
    from streammos_b200 import MultiScaleDeformableAttention as M
    value = torch.zeros(3, 16, 2, 4)
    shapes = torch.tensor([[4, 4]])
    lsi = torch.tensor([0])
    loc = torch.zeros(3, 5, 2, 1, 2, 2)
    attn = torch.zeros(3, 5, 2, 1, 2)
    with pytest.raises(RuntimeError, match="CUDA"):
        M.ms_deform_attn_forward(value, shapes, lsi, loc, attn, 256)
    with pytest.raises(AssertionError):  # batch 3 not divisible by min(3, 2): ms_deform_attn_cuda.cu:52
        M.ms_deform_attn_forward(value, shapes, lsi, loc, attn, 2)


def test_voting_rejects_cpu_and_wrong_dtype():
    from streammos_b200 import voting
    with pytest.raises(RuntimeError):
        voting.determine_voxel_labels(torch.zeros(4, 3, dtype=torch.int64), torch.zeros(4, dtype=torch.int64),
                                      (4, 4, 4), num_classes=3)


def test_dropin_registers_reference_module_names():
    from streammos_b200 import dropin
    saved = {k: sys.modules.get(k) for k in ("deep_point", "point_deep", "point_deep.cuda_kernel",
                                             "point_deep.cpu_kernel", "MultiScaleDeformableAttention")}
    try:
        dropin.install()
        import deep_point
        import point_deep.cuda_kernel as ck
        import MultiScaleDeformableAttention as MSDA
        assert hasattr(deep_point, "VoxelMaxPool") and hasattr(deep_point, "VoxelMaxPoolFunction")
        assert hasattr(ck, "voxel_maxpooling_forward") and hasattr(ck, "voxel_maxpooling_backward")
        assert hasattr(MSDA, "ms_deform_attn_forward") and hasattr(MSDA, "ms_deform_attn_backward")
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v


def test_dropin_replaces_backbone_classes_and_stem_keeps_reference_layout():
    import types
    from streammos_b200 import backbone as b200_backbone
    from streammos_b200 import dropin
    saved = {k: sys.modules.get(k) for k in ("deep_point", "point_deep", "point_deep.cuda_kernel", "point_deep.cpu_kernel",
                                             "MultiScaleDeformableAttention", "networks.backbone", "networks")}
    fake = types.ModuleType("networks.backbone")
    fake.BilinearSample = fake.PointNetStacker = object
    sys.modules["networks.backbone"] = fake
    try:
        dropin.install()
        assert fake.BilinearSample is b200_backbone.BilinearSample
        assert fake.PointNetStacker is b200_backbone.PointNetStacker
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    # parameter names of networks/backbone.py:199-250 (checkpoints load); CPU / training inputs run the torch layers,
    # exactly the reference's own code path for this module
    m = b200_backbone.PointNetStacker(7, 64, pre_bn=True, stack_num=2)
    keys = set(m.state_dict().keys())
    assert {"layer.0.layer.0.running_mean", "layer.0.layer.1.weight", "layer.0.layer.2.weight",
            "layer.1.layer.0.weight", "layer.1.layer.1.running_var"} <= keys and len(keys) == 17
    y = m.eval()(torch.randn(2, 7, 50, 1))
    assert y.shape == (2, 64, 50, 1) and float(y.min()) >= 0.0
    assert b200_backbone.PointNetStacker(7, 32, stack_num=1)(torch.randn(1, 7, 5, 1)).shape == (1, 32, 5, 1)


def test_bilinear_sample_signature():
    from streammos_b200.backbone import BilinearSample
    m = BilinearSample(in_dim=4, scale_rate=(0.5, 0.5))
    assert m.scale_rate == (0.5, 0.5)
    with pytest.raises(RuntimeError):
        m(torch.zeros(1, 2, 4, 4), torch.zeros(1, 3, 2, 1))


def test_flat_batches_pack_into_one_buffer():
    """Raw / loader batches packed into one flat buffer: same tensors, 256-byte aligned fields, one copy moves all."""
    from streammos_b200 import stream
    for make in (stream.make_host_raw_scan, stream.make_host_loader_scan):
        a = make(3, 2000, pin=False)
        p = a.pack()
        assert p._flat is not None and p.nbytes() == p._flat.numel() >= a.nbytes()
        for f in a.FIELDS:
            assert torch.equal(getattr(a, f), getattr(p, f)) and getattr(p, f).data_ptr() % 256 == p._flat.data_ptr() % 256
        q = make(4, 2000, pin=False).pack()
        q.copy_from(p)
        assert torch.equal(q._flat, p._flat)
    raw, lo = stream.make_host_raw_scan(3, 2000, pin=False), stream.make_host_loader_scan(3, 2000, pin=False)
    assert torch.equal(raw.points[0].t(), lo.pcds_xyzi[0, :4, :, 0])          # same scan in both forms
    assert torch.equal(raw.sphere_cur, lo.pcds_sphere_coord[:1]) and raw.nbytes() < lo.nbytes()
    assert not lo.coord_bev.is_contiguous() and lo.coord_bev.shape == (3, 2000, 2, 1)


def test_gpu_numa_affinity_helpers(tmp_path):
    """bench.py pins each rank to the cores next to its GPU (multi.pin_rank_to_gpu): the sysfs parsing and the
    per-rank split of a node's cores."""
    from streammos_b200 import multi
    assert multi.parse_cpulist("0-3,8,10-11\n") == [0, 1, 2, 3, 8, 10, 11]
    assert multi.parse_cpulist("") == []
    d = tmp_path / "0000:1b:00.0"
    d.mkdir()
    (d / "local_cpulist").write_text("0-27,56-83\n")
    cpus = multi.gpu_local_cpus("0000:1B:00.0", sysfs=str(tmp_path))
    assert cpus == list(range(0, 28)) + list(range(56, 84))
    assert multi.gpu_local_cpus("0000:ff:00.0", sysfs=str(tmp_path)) == []
    shares = [multi.share_of_cpus(cpus, r, 4) for r in range(4)]
    assert sorted(c for s in shares for c in s) == cpus and all(shares)     # disjoint, complete, never empty
    assert multi.share_of_cpus([5], 3, 4) == [5]
    assert multi.pin_rank_to_gpu(0, 1) is None                                 # single rank: nothing changes


def test_resident_window_stream_is_what_the_loader_would_build():
    """stream.make_host_resident_stream: the RawBatch (host-aligned frames, numpy restatement of the loader steps) and the
    ResidentRawBatch (raw scans + poses, aligned on the device in the product) describe the same model input — checked
    here against the C oracle of the loader steps, on the CPU."""
    import numpy as np
    from oracle import oracle as O
    from streammos_b200 import stream, synthetic
    n, scans = 120000, 3
    resident, aligned = stream.make_host_resident_stream(2, scans, n, pin=False)
    for i in range(scans):
        frames = []
        for k in range(3):
            r = resident[(i - k) % scans]
            raw = r.raw.numpy()[: int(r.meta[0])]
            pose = np.vstack([resident[i].poses[k].numpy().reshape(3, 4), [0, 0, 0, 1]])
            frames.append((raw, pose))
        out, cnt, _ = O.ingest_frames(frames, synthetic.RANGE_X, synthetic.RANGE_Y, synthetic.RANGE_Z, n)
        assert (cnt < n).all() and (cnt > 0.8 * n).all()
        assert np.array_equal(out.view(np.uint32), aligned[i].points.numpy().view(np.uint32))
        assert resident[i].nbytes() < 0.6 * aligned[i].nbytes()  # what crosses PCIe per scan
    linked = stream.link_window(list(resident))
    assert linked[0].older == (resident[2], resident[1]) and linked[2].older == (resident[1], resident[0])


def test_sphere_quantize_host_side():
    """ops.sphere_constants forms the four float32 constants of utils.SphereQuantize (datasets/utils.py:173-180) as numpy
    does; the oracle's restatement is within 1 ulp of the angle of a plain numpy float32 evaluation of the reference's
    lines on this host; the stream variant without host range-view coordinates carries exactly 8 bytes per point less."""
    import numpy as np
    from oracle import oracle as O
    from streammos_b200 import ops, stream, synthetic
    phi_hi, theta_hi, dphi, dtheta = ops.sphere_constants((-180.0, 180.0), synthetic.RV_THETA, synthetic.RV_SHAPE)
    assert phi_hi == float(np.float32(np.pi)) and theta_hi == float(np.float32(3.0 * np.pi / 180.0))
    assert dphi == float(np.float32(2 * np.pi / 2048)) and dtheta == float(np.float32((28.0 * np.pi / 180.0) / 64))
    pts = synthetic.make_scan(5, 20000, 1)["xyzi"][0]
    x, y, z = pts[:, 0], pts[:, 1], pts[:, 2]
    d = np.sqrt(x ** 2 + y ** 2 + z ** 2) + 1e-12                      # the reference's lines, numpy float32
    host = np.stack(((np.float32(theta_hi) - np.arcsin(z / d)) / np.float32(dtheta),
                     (np.float32(phi_hi) - np.arctan2(x, y)) / np.float32(dphi)), -1)
    assert host.dtype == np.float32
    got = O.sphere_quantize(pts, theta_range=synthetic.RV_THETA, size=synthetic.RV_SHAPE)
    diff = np.abs(got.astype(np.float64) - host)
    assert diff[:, 0].max() <= 5e-5 and diff[:, 1].max() <= 5e-4 and (diff == 0).mean() > 0.75
    full, _ = stream.make_host_resident_stream(2, 1, 120000, pin=False)
    bare, _ = stream.make_host_resident_stream(2, 1, 120000, pin=False, device_sphere=True)
    assert isinstance(bare[0], stream.ResidentScanBatch) and bare[0].coord_rv is None
    assert full[0].nbytes() - bare[0].nbytes() == 120000 * 8
    assert torch.equal(full[0].raw, bare[0].raw) and torch.equal(full[0].poses, bare[0].poses)
