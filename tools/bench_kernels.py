"""Per-kernel CUDA-event timings of the hot-path stages on one synthetic scan (explains bench.py's `value`).

    python tools/bench_kernels.py [--iters 40]

Inputs rotate over 4 scans (~100 MB each) so every call starts with a cold L2 for its big operands. Prints one
line per stage: name, µs, algorithmic MB, GB/s."""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from streammos_b200 import deep_point, ops, stream, synthetic, voting  # noqa: E402
from streammos_b200 import MultiScaleDeformableAttention as MSDA  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--iters", type=int, default=40)
ap.add_argument("--points", type=int, default=120000)
ap.add_argument("--only", default="")
a = ap.parse_args()
dev = torch.device("cuda:0")
N = a.points
hot = stream.HotPath(dev, N, seed=0, branches=False)
scans = [stream.make_host_scan(i, N).to(dev) for i in range(4)]
torch.cuda.synchronize()
rows = []


def timeit(name, fn, mb=0.0):
    if a.only and a.only not in name:
        return
    for i in range(3):
        fn(i)
    torch.cuda.synchronize()
    e0 = [torch.cuda.Event(enable_timing=True) for _ in range(a.iters)]
    e1 = [torch.cuda.Event(enable_timing=True) for _ in range(a.iters)]
    torch.cuda._sleep(30_000_000)  # ~15 ms of GPU spin: the host queues every launch below before the GPU starts them
    for i in range(a.iters):
        e0[i].record()
        fn(i)
        e1[i].record()
    torch.cuda.synchronize()
    ts = sorted(x.elapsed_time(y) * 1e3 for x, y in zip(e0, e1))
    med = ts[len(ts) // 2]
    rows.append((name, med, mb))
    print("%-34s %8.1f us %8.1f MB %8.0f GB/s" % (name, med, mb, mb / med * 1e3 if mb else 0), flush=True)


with torch.no_grad():
    S = lambda i: scans[i % 4]
    specs = lambda b: [(b.coord_bev, (512, 512), (1.0, 1.0)), (b.coord_rv, (32, 1024), (0.5, 0.5)),
                       (b.coord_bev[:1], (256, 256), (0.5, 0.5)), (b.coord_rv, (16, 512), (0.25, 0.25)),
                       (b.coord_bev[:1], (128, 128), (0.25, 0.25))]
    timeit("plan x5 (4 kernels)", lambda i: ops.pool_plan_multi(specs(S(i))))
    timeit("plan x5 + gather taps x4", lambda i: ops.pool_plan_multi(specs(S(i)), gather_taps=[False, True, True, True, True]))
    plans = [ops.pool_plan_multi(specs(s)) for s in scans]
    tplans = [ops.pool_plan_multi(specs(s), gather_taps=[False, True, True, True, True]) for s in scans]
    ws = ops.pool_workspace(3, 64, N, dev)
    out1 = torch.empty(3, 64, 512, 512, device=dev)
    for s, p in zip(scans, plans):
        ops.voxel_maxpool_forward(s.feat, p[0], out=out1, workspace=ws)
    mb_in = 4 * 3 * 64 * N / 1e6
    timeit("pool1 permute", lambda i: ops.voxel_maxpool_forward(S(i).feat, plans[i % 4][0], out=out1, stages=1 | 8, workspace=ws), mb_in)
    timeit("pool1 permute+reduce", lambda i: ops.voxel_maxpool_forward(S(i).feat, plans[i % 4][0], out=out1, stages=1, workspace=ws), mb_in)
    timeit("pool1 combine", lambda i: ops.voxel_maxpool_forward(S(i).feat, plans[i % 4][0], out=out1, stages=2, workspace=ws))
    timeit("pool1 write", lambda i: ops.voxel_maxpool_forward(S(i).feat, plans[i % 4][0], out=out1, stages=4, workspace=ws), 4 * 3 * 64 * 512 * 512 / 1e6)
    timeit("pool1 all stages", lambda i: ops.voxel_maxpool_forward(S(i).feat, plans[i % 4][0], out=out1, workspace=ws), mb_in + 4 * 3 * 64 * 512 * 512 / 1e6)
    # gathers: scan order vs cell order, NCHW grids, point-major rows out
    cfg = [("gather1 x0 32ch@256^2", hot.x0, hot.g_half, "bev", 2), ("gather3 x1 64ch@128^2", hot.x1, hot.g_quarter, "bev", 4),
           ("gather5 dec 64ch@256^2", hot.dec, hot.g_half, "bev", 2)]
    for name, grid, mod, view, pi in cfg:
        C, H, W = grid.shape[1:]
        mb = (4 * C * H * W + 8 * N + 4 * C * N) / 1e6
        timeit(name + " scan-order", lambda i: mod(grid, S(i).coord_bev[:1]), mb)
        timeit(name + " cell-order", lambda i: mod(grid, S(i).coord_bev[:1], plans[i % 4][pi]), mb)
        timeit(name + " plan taps", lambda i: mod(grid, S(i).coord_bev[:1], tplans[i % 4][pi]), mb)
    x0_pt = [hot.g_half(hot.x0, s.coord_bev[:1]) for s in scans]
    x0_rv = deep_point.VoxelMaxPool(x0_pt[0], scans[0].coord_rv, (32, 1024), (0.5, 0.5), plans[0][1])
    mb = (4 * 32 * 32 * 1024 + 8 * N + 4 * 32 * N) / 1e6
    timeit("gather2 rv 32ch@32x1024 scan-order", lambda i: hot.g_half(x0_rv, S(i).coord_rv), mb)
    timeit("gather2 rv 32ch@32x1024 cell-order", lambda i: hot.g_half(x0_rv, S(i).coord_rv, plans[i % 4][1]), mb)
    timeit("gather2 rv 32ch@32x1024 plan taps", lambda i: hot.g_half(x0_rv, S(i).coord_rv, tplans[i % 4][1]), mb)
    # small pools (point-major input): stages
    for name, C, pi, size in (("pool2 rv32x1024 c32", 32, 1, (32, 1024)), ("pool3 bev256 c32", 32, 2, (256, 256))):
        o = torch.empty(1, C, *size, device=dev)
        w2 = ops.pool_workspace(1, C, N, dev)
        f = x0_pt
        mb = (4 * C * N + 8 * N + 4 * C * size[0] * size[1]) / 1e6
        timeit(name + " reduce", lambda i: ops.voxel_maxpool_forward(f[i % 4], plans[i % 4][pi], out=o, stages=1, workspace=w2))
        timeit(name + " combine", lambda i: ops.voxel_maxpool_forward(f[i % 4], plans[i % 4][pi], out=o, stages=2, workspace=w2))
        timeit(name + " write", lambda i: ops.voxel_maxpool_forward(f[i % 4], plans[i % 4][pi], out=o, stages=4, workspace=w2))
        timeit(name + " all", lambda i: ops.voxel_maxpool_forward(f[i % 4], plans[i % 4][pi], out=o, workspace=w2), mb)
    x1_pt = [hot.g_quarter(hot.x1, s.coord_bev[:1]) for s in scans]
    for name, C, pi, size in (("pool4 rv16x512 c64", 64, 3, (16, 512)), ("pool5 bev128 c64", 64, 4, (128, 128))):
        o = torch.empty(1, C, *size, device=dev)
        w2 = ops.pool_workspace(1, C, N, dev)
        mb = (4 * C * N + 8 * N + 4 * C * size[0] * size[1]) / 1e6
        timeit(name + " reduce", lambda i: ops.voxel_maxpool_forward(x1_pt[i % 4], plans[i % 4][pi], out=o, stages=1, workspace=w2))
        timeit(name + " combine", lambda i: ops.voxel_maxpool_forward(x1_pt[i % 4], plans[i % 4][pi], out=o, stages=2, workspace=w2))
        timeit(name + " write", lambda i: ops.voxel_maxpool_forward(x1_pt[i % 4], plans[i % 4][pi], out=o, stages=4, workspace=w2))
        timeit(name + " all", lambda i: ops.voxel_maxpool_forward(x1_pt[i % 4], plans[i % 4][pi], out=o, workspace=w2), mb)
    value = hot.memory.view(1, 4096, 4, 32)
    timeit("msda fwd (one layer)", lambda i: MSDA.ms_deform_attn_forward(value, hot.shapes, hot.lsi, S(i).loc[0], S(i).attn[0], 256), 4.98)
    # voting pieces (reference API)
    pts = hot.local_pts.view(-1, 4)
    P = pts.shape[0]
    timeit("vote quantize", lambda i: voting.Quantize(pts, synthetic.RANGE_X, synthetic.RANGE_Y, synthetic.RANGE_Z, hot.size), 28 * P / 1e6)
    q = voting.Quantize(pts, synthetic.RANGE_X, synthetic.RANGE_Y, synthetic.RANGE_Z, hot.size)
    timeit("vote coords .to(int64) [torch]", lambda i: q.to(torch.int64), 36 * P / 1e6)
    timeit("vote labels .to(int64) [torch]", lambda i: hot.local_pred.view(-1).to(torch.int64), 9 * P / 1e6)
    coords = q.to(torch.int64)
    labels = hot.local_pred.view(-1).to(torch.int64)
    timeit("vote determine_voxel_labels", lambda i: voting.determine_voxel_labels(coords, labels, hot.size, num_classes=3), (32 * P + 8 * 512 * 512 * 30) / 1e6)
    vl = voting.determine_voxel_labels(coords, labels, hot.size, num_classes=3)
    timeit("vote point labels", lambda i: voting.get_point_labels_from_voxel_labels(coords[8 * N:], vl, hot.size), 40 * N / 1e6)
    timeit("vote fused API", lambda i: ops.vote_fused(pts, hot.local_pred.view(-1), N, hot.mins, hot.deltas, hot.size, 3))
    timeit("instance votes (32 boxes)", lambda i: ops.instance_vote(pts, labels, hot.box_lo, hot.box_hi), 24 * P / 1e6)
    lscans = [stream.make_host_loader_scan(i, N).to(dev) for i in range(4)]
    timeit("point stem 3x7xN -> 3x64xN (fused)", lambda i: hot.point_pre(lscans[i % 4].pcds_xyzi), (4 * 3 * 7 * N + 4 * 3 * 64 * N) / 1e6)
    timeit("point stem (torch layers)", lambda i: hot.stem.layer(lscans[i % 4].pcds_xyzi), (4 * 3 * 7 * N + 4 * 3 * 64 * N) / 1e6)
    timeit("whole step (eager, serial)", lambda i: hot.step(S(i)))
    # instance clustering: cluster() of voxel_instance_voting.py on a scan with ~30 moving objects
    import time
    import numpy as np
    rng = np.random.default_rng(0)
    for n_obj, per_obj in ((30, 100), (60, 400)):
        objs = [np.array([rng.uniform(-40, 40), rng.uniform(-40, 40), rng.uniform(-1.5, -0.5)]) +
                rng.uniform(-1, 1, (per_obj, 3)) * np.array([1.0, 0.5, 0.4]) for _ in range(n_obj)]
        fgp = np.concatenate(objs + [np.stack([rng.uniform(-50, 50, 300), rng.uniform(-50, 50, 300),
                                               rng.uniform(-3, 1, 300)], -1)]).astype(np.float32)
        M = len(fgp)
        cur = pts[8 * N:].clone()
        idx = torch.from_numpy(rng.permutation(N)[:M]).to(dev)
        cur[idx, :3] = torch.from_numpy(fgp).to(dev)
        bf = torch.zeros(N, dtype=torch.int32, device=dev)
        bf[idx] = 2
        pred = torch.randint(0, 3, (N,), device=dev)
        tag = "cluster M=%d" % M
        timeit(tag + " dbscan+boxes (9 kernels)", lambda i: ops.cluster_boxes(cur, bf))
        timeit(tag + " cluster() whole", lambda i: voting.cluster(cur, pred, bf, pts, labels))
        try:
            from sklearn.cluster import DBSCAN
            x = cur[bf == 2][:, :3].cpu().numpy()
            t0 = time.perf_counter()
            ref = DBSCAN(eps=0.3, min_samples=5).fit_predict(x)
            dt = time.perf_counter() - t0
            same = np.array_equal(ref, ops.cluster_boxes(cur, bf)["fg_label"][:M].cpu().numpy())
            print("%-34s %8.1f us   (sklearn DBSCAN alone on the host, %d clusters, labels identical: %s)"
                  % (tag + " [host]", dt * 1e6, ref.max() + 1, same), flush=True)
        except ImportError:
            pass
    # training-side kernels: not on the streaming path, timed here for the record
    go1 = torch.randn(3, 64, 512, 512, device=dev)
    mb_b1 = (4 * 3 * 64 * 512 * 512 + 4 * 3 * 64 * N * 2 + 8 * 3 * N) / 1e6
    timeit("bwd pool1 (3x64xN <- 3x64x512^2)", lambda i: ops.voxel_maxpool_backward(S(i).feat, plans[i % 4][0], out1, go1), mb_b1)
    gp = torch.randn(1, 64, N, 1, device=dev)
    timeit("bwd gather5 64ch@256^2", lambda i: ops.bilinear_gather_backward(gp, S(i).coord_bev[:1], (0.5, 0.5), 256, 256),
           (4 * 64 * 256 * 256 + 8 * N + 4 * 64 * N) / 1e6)
    gq = torch.randn(1, 64, N, 1, device=dev)
    timeit("bwd gather3 64ch@128^2", lambda i: ops.bilinear_gather_backward(gq, S(i).coord_bev[:1], (0.25, 0.25), 128, 128),
           (4 * 64 * 128 * 128 + 8 * N + 4 * 64 * N) / 1e6)
    gout = torch.randn(1, 4096, 128, device=dev)
    timeit("bwd msda (one layer)", lambda i: MSDA.ms_deform_attn_backward(value, hot.shapes, hot.lsi, S(i).loc[0], S(i).attn[0], gout, 256), 12.3)
