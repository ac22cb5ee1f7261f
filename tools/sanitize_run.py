"""One small invocation of EVERY kernel of libstreammos_b200.so (checked against the oracle), sized to finish in a
minute under compute-sanitizer:

    compute-sanitizer --tool memcheck  --error-exitcode 1 python tools/sanitize_run.py
    compute-sanitizer --tool racecheck --error-exitcode 1 python tools/sanitize_run.py

SURVEY §5: "new kernels must be compute-sanitizer --tool racecheck clean". On this pool compute-sanitizer is CLOSED
(gpurun answers "compute-sanitizer is closed on this pool and stays closed", profiles/r2_compute_sanitizer_closed.txt), so
the script is run plain — every result compared with the oracle — and the race / bounds arguments are made by construction
in DESIGN.md §4.6.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402
from streammos_b200 import MultiScaleDeformableAttention as MSDA  # noqa: E402
from streammos_b200 import deep_point, ops, plan_cache, stream, synthetic, voting  # noqa: E402
from streammos_b200.backbone import BilinearSample, PointNetStacker  # noqa: E402

dev = torch.device("cuda:0")
rng = np.random.default_rng(0)
t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
done = []


def ok(name):
    torch.cuda.synchronize()
    done.append(name)
    print("ok", name, flush=True)


# ---- pooling: plans (single, batched, cached + prefetched), forward on every input layout, backward -------------------
B, C, N, H, W = 2, 32, 6000, 48, 64
ind = np.stack([rng.uniform(-2, H * 2 + 4, (B, N)), rng.uniform(-2, W * 2 + 4, (B, N))], -1).astype(np.float32)[..., None]
ind[:, ::17] = -1000.0
feat = rng.standard_normal((B, C, N, 1)).astype(np.float32)
want = O.voxel_maxpool_forward(feat, ind, (H, W), (0.5, 0.5))
for layout in ("channel_major", "point_major"):
    f = t(feat)
    if layout == "point_major":
        f = f.permute(0, 2, 3, 1).contiguous().permute(0, 3, 1, 2)
    for env in ("1", "0"):
        os.environ["SMOS_PERM_LDG"] = env
        plan_cache.clear()
        got = deep_point.VoxelMaxPool(f, t(ind), (H, W), (0.5, 0.5))
        assert np.array_equal(got.cpu().numpy(), want), (layout, env)
os.environ.pop("SMOS_PERM_LDG")
ok("VoxelMaxPool forward (LDG / TMA permute, point-major reduce, combine, write)")
plans = ops.pool_plan_multi([(t(ind), (H, W), (0.5, 0.5)), (t(ind), (24, 32), (0.25, 0.25))], gather_taps=[True, True])
g3 = t(rng.standard_normal((B, C, H, W)).astype(np.float32))
assert np.array_equal(ops.voxel_maxpool_forward(t(feat), plans[0]).cpu().numpy(), want)
fx = t(feat).requires_grad_(True)
out = deep_point.VoxelMaxPool(fx, t(ind), (H, W), (0.5, 0.5))
out.backward(g3)
assert np.array_equal(fx.grad.cpu().numpy()[..., 0], O.voxel_maxpool_backward(feat, ind, want, g3.cpu().numpy(), (0.5, 0.5)))
ok("batched plans with gather records, pooling backward")
from streammos_b200.point_deep import cuda_kernel  # noqa: E402
vo = torch.zeros(B, C, H, W, device=dev)
idx = torch.full((B, N), -1, dtype=torch.int64, device=dev)
sz, sd = torch.tensor([H, W], device=dev), torch.tensor([W, 1], device=dev)
cuda_kernel.voxel_maxpooling_forward(t(feat), t(ind), vo, idx, sz, sd, sz, torch.tensor([0.5, 0.5], device=dev))
assert np.array_equal(vo.cpu().numpy(), want)
ok("point_deep.cuda_kernel lower boundary (device-resident scales)")

# ---- gathers: planar (scan order, cell order, plan records), channels-last, non-dense strides, backward --------------
grid = rng.standard_normal((B, C, H, W)).astype(np.float32)
wantg = O.bilinear_sample(grid, ind, (0.5, 0.5))
for cl in (False, True):
    gt = t(grid).contiguous(memory_format=torch.channels_last) if cl else t(grid)
    for pm in (False, True):
        for order in (None, plans[0], "auto"):
            m = BilinearSample(C, (0.5, 0.5))
            m.point_major_out = pm
            m.auto_order = order == "auto"
            got = m(gt, t(ind), None if order == "auto" else order)
            np.testing.assert_allclose(got[..., 0].cpu().numpy(), wantg, rtol=1e-5, atol=1e-6)
gs = t(grid)[:, :, :, ::2]  # non-dense planes
np.testing.assert_allclose(BilinearSample(C, (0.5, 0.25))(gs, t(ind))[..., 0].cpu().numpy(),
                           O.bilinear_sample(grid[:, :, :, ::2], ind, (0.5, 0.25)), rtol=1e-5, atol=1e-6)
gg = t(grid).requires_grad_(True)
BilinearSample(C, (0.5, 0.5))(gg, t(ind)).backward(t(rng.standard_normal((B, C, N, 1)).astype(np.float32)))
ok("BilinearSample forward (all variants) + backward")

# ---- deformable attention fwd / bwd, fp32 + fp64, vector and scalar channel counts --------------------------------------
for D, dt in ((32, torch.float32), (30, torch.float64), (71, torch.float32)):
    S_, M_, P_ = 12 * 12, 2, 4
    val = t(rng.standard_normal((1, S_, M_, D))).to(dt).requires_grad_(True)
    loc = t(rng.uniform(-0.1, 1.1, (1, S_, M_, 1, P_, 2))).to(dt).requires_grad_(True)
    att = t(rng.uniform(0, 1, (1, S_, M_, 1, P_))).to(dt).requires_grad_(True)
    shp, lsi = torch.tensor([[12, 12]], device=dev), torch.zeros(1, dtype=torch.int64, device=dev)
    from streammos_b200.functions import MSDeformAttnFunction  # noqa: E402
    o = MSDeformAttnFunction.apply(val, shp, lsi, loc, att, 64)
    o.sum().backward()
    ref = O.ms_deform_attn_forward(val.detach().cpu().numpy(), shp.cpu().numpy(), lsi.cpu().numpy(), loc.detach().cpu().numpy(),
                                   att.detach().cpu().numpy())
    np.testing.assert_allclose(o.detach().cpu().numpy(), ref, rtol=1e-4, atol=1e-5)
ok("MSDeformAttn forward + backward")

# ---- voting: staging, int64 API, fused API, streaming API, instance votes, memory ring ---------------------------------
S, n = 4, 5000
ring_p = np.concatenate([rng.uniform(-52, 52, (S, n, 2)), rng.uniform(-4.5, 2.5, (S, n, 1)), rng.uniform(0, 1, (S, n, 1))], -1).astype(np.float32)
ring_l = rng.integers(0, 3, (S, n)).astype(np.uint8)
dp, dl = t(ring_p), t(ring_l)
q, coords, labels = voting.quantize_staged(dp, dl, (-50.0, 50.0), (-50.0, 50.0), (-4.0, 2.0), (512, 512, 30),
                                           new_points=t(ring_p[0] * 0.5), new_pred=t(ring_l[1]), cur_slot=3, hist_slot=2)
vl = voting.determine_voxel_labels(coords, labels, (512, 512, 30), num_classes=3)
assert np.array_equal(vl.cpu().numpy(), O.determine_voxel_labels(coords.cpu().numpy(), labels.cpu().numpy(), (512, 512, 30), 3))
pl = voting.get_point_labels_from_voxel_labels(coords[3 * n:], vl, (512, 512, 30))
vl5 = voting.determine_voxel_labels(coords, labels + 2, (512, 512, 30), num_classes=5)
ops.memory_push(dp[0], dl[0], dp[3], dl[3], dp[1], dl[1])
vf, pf = ops.vote_fused(dp.view(-1, 4), dl.view(-1), n, (-50.0, -50.0, -4.0), (100 / 512, 100 / 512, 6 / 30), (512, 512, 30), 3)
lo = np.concatenate([rng.uniform(-45, 40, (40, 2)), rng.uniform(-3, 0, (40, 1))], 1).astype(np.float32)
hi = lo + rng.uniform(1, 5, (40, 3)).astype(np.float32)
s1 = ops.instance_vote(dp.view(-1, 4), labels, t(lo), t(hi))
ws = ops.instance_vote_workspace(40, dev)
s2 = ops.instance_vote(dp.view(-1, 4), labels, t(lo), t(hi), workspace=ws)
assert torch.equal(s1, s2)
voter = voting.StreamingVoter()
for k in range(3):
    pose = np.eye(4)
    pose[0, 3] = 0.3 * k
    voter.push(dp[k], dl[k], pose)
voter.vote()
ok("voting (staging, int64 API packed + unpacked, fused, streaming, instance votes with / without workspace, ring push)")

# ---- instance clustering ------------------------------------------------------------------------------------------------
M = 1500
blobs = np.concatenate([rng.normal(c, 0.15, (150, 3)) for c in rng.uniform(-20, 20, (8, 3))] + [rng.uniform(-30, 30, (300, 3))])
pts = np.concatenate([blobs, np.zeros((len(blobs), 1))], 1).astype(np.float32)
st = ops.cluster_boxes(t(pts), torch.full((len(pts),), 2, dtype=torch.int32, device=dev))
assert np.array_equal(st["fg_label"][:len(pts)].cpu().numpy(), O.dbscan(pts[:, :3].astype(np.float64), 0.3, 5))
pred = torch.randint(0, 3, (len(pts),), device=dev)
voting.cluster(t(pts), pred, torch.full((len(pts),), 2, dtype=torch.int32, device=dev), dp.view(-1, 4), labels)
ok("instance clustering (DBSCAN + boxes + vote + write-back)")

# ---- PointNet stem + form_batch -----------------------------------------------------------------------------------------
stem = PointNetStacker(7, 64, pre_bn=True, stack_num=2).eval().to(dev)
raw = t(synthetic.make_scan(3, 4096, 3)["xyzi"])
with torch.no_grad():
    f7, coord = ops.form_batch(raw, synthetic.RANGE_X, synthetic.RANGE_Y, synthetic.RANGE_Z, synthetic.BEV_SHAPE)
    a_ = stem(f7)
    b_, _ = ops.point_stem_forward_raw(raw, synthetic.RANGE_X, synthetic.RANGE_Y, synthetic.RANGE_Z, synthetic.BEV_SHAPE,
                                       *stem.fused_parameters())
    assert torch.equal(a_, b_)
    os.environ["SMOS_STEM_TC"] = "1"
    stem(f7)
    os.environ.pop("SMOS_STEM_TC")
ok("form_batch + PointNet stem (FMA and tensor-core variants, raw-scan variant)")

# ---- one whole scan through the harness (reference signatures, plan cache, in-place memory) ----------------------------
hot = stream.HotPath(dev, n_points=8192, seed=1)
with torch.no_grad():
    for i in range(3):
        hot.step(stream.make_host_scan(i, 8192).to(dev))
ok("HotPath.step x3")
print("ALL OK:", len(done), "groups")
