"""Generate tests/golden/*.npz by running the REFERENCE ITSELF (read-only tree at /root/reference) on
seeded synthetic inputs, in the build container. The vectors pin oracle/ (and through it the CUDA
kernels); the reference cannot travel to the GPU box, the vectors do.

What runs, unmodified:
  * deep_point/__init__.py on top of deep_point/src/point_deep.cpp compiled by oracle/build_ref.py
  * networks/backbone.py:BilinearSample                       (loaded by file path)
  * deformattn/functions/ms_deform_attn_func.py:ms_deform_attn_core_pytorch (+ torch autograd for grads)
  * voxel_voting.py / voxel_instance_voting.py functions, extracted with `ast` because the scripts run
    argparse at import time (voxel_voting.py:128-136)
  * networks/backbone.py:PointNetStacker(7, 64, pre_bn=True, stack_num=2).eval()   (the stem of models/StreamMOS.py:77)
  * datasets/utils.py Quantize / SphereQuantize + datasets/data_StreamMOS.py make_point_feat (the loader's form_batch;
    SphereQuantize's own output: sphere_a)

    python tools/make_golden.py [--ref /root/reference]
"""
import argparse
import ast
import importlib.util
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
GOLD = os.path.join(ROOT, "tests", "golden")


def load_by_path(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def extract_functions(path, names, glob):
    """exec only the named top-level function definitions of a script."""
    tree = ast.parse(open(path).read())
    body = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in names]
    found = {n.name for n in body}
    missing = set(names) - found
    assert not missing, missing
    exec(compile(ast.Module(body=body, type_ignores=[]), path, "exec"), glob)
    return glob


def ref_deep_point(ref):
    from oracle import build_ref
    build_ref.build(ref)
    cpu = build_ref.load()
    assert cpu is not None
    pkg = types.ModuleType("point_deep")
    stub = types.ModuleType("point_deep.cuda_kernel")  # never called: CPU tensors only
    pkg.cpu_kernel, pkg.cuda_kernel = cpu, stub
    sys.modules["point_deep"] = pkg
    sys.modules["point_deep.cpu_kernel"] = cpu
    sys.modules["point_deep.cuda_kernel"] = stub
    return load_by_path("ref_deep_point", os.path.join(ref, "deep_point", "__init__.py"))


def synth_coords(rng, B, N, H, W, scale, frac_edge=0.05):
    """coords such that coord*scale spans the grid, with a few in (-1,0), >= size and far pads."""
    c = np.stack([rng.uniform(0, H / scale[0], (B, N)), rng.uniform(0, W / scale[1], (B, N))], -1)
    k = max(1, int(N * frac_edge))
    c[:, :k, 0] = rng.uniform(-0.999, 0, (B, k)) / scale[0]            # truncates to cell 0 (valid)
    c[:, k:2 * k, 1] = W / scale[1] + rng.uniform(0, 3, (B, k))        # out of range high
    c[:, 2 * k:3 * k] = -1000.0                                        # pads (data_StreamMOS.py:570)
    c[:, 3 * k:4 * k, 0] = rng.uniform(-5, -1.001, (B, k)) / scale[0]  # out of range low
    return c.astype(np.float32)


def gen_pool(ref, rng):
    dp = ref_deep_point(ref)
    cases = {}
    for name, (B, C, N, H, W, scale) in {
        "pool_a": (2, 5, 4000, 24, 40, (0.5, 0.25)),
        "pool_b": (3, 8, 3000, 64, 64, (1.0, 1.0)),
        "pool_c": (1, 33, 2500, 16, 96, (0.25, 0.5)),
    }.items():
        feat = rng.standard_normal((B, C, N, 1)).astype(np.float32)
        feat[:, :, ::7] = np.round(feat[:, :, ::7], 1)  # force exact ties
        ind = synth_coords(rng, B, N, H, W, scale)[..., None]
        tf = torch.from_numpy(feat).requires_grad_(True)
        out = dp.VoxelMaxPool(tf, torch.from_numpy(ind), (H, W), scale)
        gout = rng.standard_normal(out.shape).astype(np.float32)
        out.backward(torch.from_numpy(gout))
        cases[name] = dict(feat=feat, ind=ind, H=H, W=W, scale=np.array(scale, np.float32),
                           out=out.detach().numpy(), gout=gout, gfeat=tf.grad.numpy())
    for k, v in cases.items():
        np.savez_compressed(os.path.join(GOLD, k + ".npz"), **v)
    return list(cases)


def gen_bilinear(ref, rng):
    bb = load_by_path("ref_backbone", os.path.join(ref, "networks", "backbone.py"))
    names = []
    for name, (B, C, H, W, N, scale) in {
        "bilinear_a": (2, 6, 16, 24, 3000, (0.5, 0.5)),
        "bilinear_b": (1, 32, 32, 128, 2000, (0.25, 0.25)),
    }.items():
        grid = rng.standard_normal((B, C, H, W)).astype(np.float32)
        coord = synth_coords(rng, B, N, H, W, scale)[..., None]
        # push a few points exactly onto / just beyond the last pixel
        coord[:, -5:, 0, 0] = (H - 1) / scale[0]
        coord[:, -3:, 1, 0] = (W - 1) / scale[1] + 0.5
        m = bb.BilinearSample(in_dim=C, scale_rate=scale)
        tg = torch.from_numpy(grid).requires_grad_(True)
        out = m(tg, torch.from_numpy(coord))
        gout = rng.standard_normal(out.shape).astype(np.float32)
        out.backward(torch.from_numpy(gout))
        out64 = m(torch.from_numpy(grid).double(), torch.from_numpy(coord).double())
        np.savez_compressed(os.path.join(GOLD, name + ".npz"), grid=grid, coord=coord,
                            scale=np.array(scale, np.float32), out=out.detach().numpy(), out64=out64.numpy(),
                            gout=gout, ggrid=tg.grad.numpy())
        names.append(name)
    return names


def gen_msda(ref, rng):
    sys.modules.setdefault("MultiScaleDeformableAttention", types.ModuleType("MultiScaleDeformableAttention"))
    fn = load_by_path("ref_msda_func", os.path.join(ref, "deformattn", "functions", "ms_deform_attn_func.py"))
    names = []

    def run(name, value, shapes, loc, attn):
        lsi = torch.cat((shapes.new_zeros((1,)), shapes.prod(1).cumsum(0)[:-1]))
        v = value.double().requires_grad_(True)
        lo = loc.double().requires_grad_(True)
        a = attn.double().requires_grad_(True)
        out = fn.ms_deform_attn_core_pytorch(v, shapes, lo, a)
        gout = torch.from_numpy(rng.standard_normal(tuple(out.shape)))
        out.backward(gout)
        out32 = fn.ms_deform_attn_core_pytorch(value.float(), shapes, loc.float(), attn.float())
        np.savez_compressed(os.path.join(GOLD, name + ".npz"), value=value.numpy(), shapes=shapes.numpy(),
                            lsi=lsi.numpy(), loc=loc.numpy(), attn=attn.numpy(), out64=out.detach().numpy(),
                            out32=out32.numpy(), gout=gout.numpy(), gvalue=v.grad.numpy(), gloc=lo.grad.numpy(),
                            gattn=a.grad.numpy())
        names.append(name)

    # the reference's own test configuration (deformattn/test.py:23-38): seed 3, N,M,D=1,2,2 Lq,L,P=2,2,2
    torch.manual_seed(3)
    N, M, D, Lq, L, P = 1, 2, 2, 2, 2, 2
    shapes = torch.as_tensor([(6, 4), (3, 2)], dtype=torch.long)
    S = int(shapes.prod(1).sum())
    value = torch.rand(N, S, M, D) * 0.01
    loc = torch.rand(N, Lq, M, L, P, 2)
    attn = torch.rand(N, Lq, M, L, P) + 1e-5
    attn /= attn.sum(-1, keepdim=True).sum(-2, keepdim=True)
    run("msda_reftest", value, shapes, loc, attn)
    # gradcheck channel counts of test.py:85 that are cheap to store
    for D in (30, 32, 71):
        value = torch.rand(N, S, M, D) * 0.01
        loc = torch.rand(N, Lq, M, L, P, 2)
        attn = torch.rand(N, Lq, M, L, P) + 1e-5
        attn /= attn.sum(-1, keepdim=True).sum(-2, keepdim=True)
        run("msda_reftest_d%d" % D, value, shapes, loc, attn)
    # StreamMOS configuration in small: 1 level, 4 heads x 32, 4 points, queries = cell centres + offsets
    Hs = Ws = 16
    B, M, D, P = 2, 4, 32, 4
    shapes = torch.as_tensor([(Hs, Ws)], dtype=torch.long)
    value = torch.randn(B, Hs * Ws, M, D)
    ys, xs = torch.meshgrid(torch.linspace(0.5, Hs - 0.5, Hs), torch.linspace(0.5, Ws - 0.5, Ws), indexing="ij")
    ref_pts = torch.stack((xs.reshape(-1) / Ws, ys.reshape(-1) / Hs), -1)  # multi_view_encoder.py:254-266
    loc = ref_pts[None, :, None, None, None, :] + torch.randn(B, Hs * Ws, M, 1, P, 2) * 2.0 / Hs
    attn = torch.softmax(torch.randn(B, Hs * Ws, M, 1 * P), -1).view(B, Hs * Ws, M, 1, P)
    run("msda_streammos_small", value, shapes, loc, attn)
    return names


def gen_msda_module(ref, rng):
    """The reference's MSDeformAttn MODULE (deformattn/modules/ms_deform_attn.py) run on the CPU: its own forward with the
    compiled sampling core replaced by the reference's ms_deform_attn_core_pytorch. Stores the module's weights, inputs,
    the tensors it hands to the core (raw offsets, logits) and its output, for the 2-d (StreamMOS) and 4-d reference
    point forms."""
    sys.modules.setdefault("MultiScaleDeformableAttention", types.ModuleType("MultiScaleDeformableAttention"))
    if ref not in sys.path:
        sys.path.insert(0, ref)
    import deformattn.modules.ms_deform_attn as mod
    import deformattn.functions.ms_deform_attn_func as fn

    class _Core:  # stands in for MSDeformAttnFunction: same positional call, the reference's CPU core
        @staticmethod
        def apply(value, shapes, lsi, loc, attn, im2col_step):
            _Core.seen = (loc.detach().clone(), attn.detach().clone())
            return fn.ms_deform_attn_core_pytorch(value, shapes, loc, attn)

    mod.MSDeformAttnFunction = _Core
    names = []
    for name, d_model, L, M, P, shapes, ref_dim, N, Lq in (
            ("msda_module_streammos", 128, 1, 4, 4, [(12, 12)], 2, 1, 96),   # CENet_Transformer: D_MODEL 128, 1 level
            ("msda_module_boxes", 64, 2, 2, 3, [(6, 4), (3, 2)], 4, 1, 9)):
        torch.manual_seed(11)
        m = mod.MSDeformAttn(d_model, L, M, P).eval()
        with torch.no_grad():  # the default init zeroes the offset / logit weights: give them life
            m.sampling_offsets.weight.copy_(torch.randn_like(m.sampling_offsets.weight) * 0.05)
            m.attention_weights.weight.copy_(torch.randn_like(m.attention_weights.weight) * 0.2)
            m.attention_weights.bias.copy_(torch.randn_like(m.attention_weights.bias) * 0.2)
        sh = torch.as_tensor(shapes, dtype=torch.long)
        lsi = torch.cat((sh.new_zeros((1,)), sh.prod(1).cumsum(0)[:-1]))
        S = int(sh.prod(1).sum())
        query = torch.randn(N, Lq, d_model)
        src = torch.randn(N, S, d_model)
        refp = torch.rand(N, Lq, L, ref_dim)
        if ref_dim == 4:
            refp[..., 2:] = refp[..., 2:] * 0.5 + 0.1
        with torch.no_grad():
            out = m(query, refp, src, sh, lsi)
            value = m.value_proj(src).view(N, S, M, d_model // M)
            offs = m.sampling_offsets(query).view(N, Lq, M, L, P, 2)
            logits = m.attention_weights(query).view(N, Lq, M, L * P)
            core = fn.ms_deform_attn_core_pytorch(value, sh, *_Core.seen)
        sd = {"w_" + k.replace(".", "__"): v.numpy() for k, v in m.state_dict().items()}
        np.savez_compressed(os.path.join(GOLD, name + ".npz"), query=query.numpy(), src=src.numpy(), ref=refp.numpy(),
                            shapes=sh.numpy(), lsi=lsi.numpy(), out=out.numpy(), value=value.numpy(), offsets=offs.numpy(),
                            logits=logits.numpy(), loc=_Core.seen[0].numpy(), attn=_Core.seen[1].numpy(),
                            core_out=core.numpy(), dims=np.array([d_model, L, M, P]), **sd)
        names.append(name)
    return names


def gen_voting(ref, rng):
    g = {"torch": torch, "np": np}
    extract_functions(os.path.join(ref, "voxel_voting.py"),
                      ["get_point_labels_from_voxel_labels", "determine_voxel_labels", "Quantize"], g)
    size = (32, 24, 10)
    rx, ry, rz = (-50.0, 50.0), (-50.0, 50.0), (-4.0, 2.0)
    P, Pc = 20000, 3000
    pts = np.stack([rng.uniform(rx[0] + 1e-3, rx[1] - 1e-3, P), rng.uniform(ry[0] + 1e-3, ry[1] - 1e-3, P),
                    rng.uniform(rz[0] + 1e-3, rz[1] - 1e-3, P), rng.uniform(0, 1, P)], -1).astype(np.float32)
    # concentrate points so that voxels collect several votes and ties occur
    pts[: P // 2, :3] = (pts[: P // 2, :3] * np.array([0.2, 0.2, 0.5], np.float32))
    labels = rng.integers(0, 3, P).astype(np.int64)
    quan = g["Quantize"](torch.from_numpy(pts), range_x=rx, range_y=ry, range_z=rz, size=size)
    coords = quan.to(torch.int64)
    vl = g["determine_voxel_labels"](coords, torch.from_numpy(labels), size)
    cur = coords[P - Pc:].clone()
    cur[:50, 0] = size[0] + 3     # out-of-range lookups -> 0
    cur[50:80, 2] = -1
    pl = g["get_point_labels_from_voxel_labels"](cur, vl, size)
    np.savez_compressed(os.path.join(GOLD, "voting_a.npz"), pts=pts, labels=labels, size=np.array(size),
                        rx=np.array(rx), ry=np.array(ry), rz=np.array(rz), quan=quan.numpy(),
                        coords=coords.numpy(), voxel_labels=vl.numpy(), cur=cur.numpy(), point_labels=pl.numpy())

    # numpy Quantize of voxel_instance_voting.py:117-135 must agree with the torch one on CPU
    g2 = {"np": np}
    extract_functions(os.path.join(ref, "voxel_instance_voting.py"), ["Quantize"], g2)
    q_np = g2["Quantize"](pts, range_x=rx, range_y=ry, range_z=rz, size=size)
    assert q_np.dtype == np.float32 and np.array_equal(q_np, quan.numpy())
    return ["voting_a"]


def gen_instance(ref, rng):
    import scipy
    from scipy.spatial import ConvexHull, Delaunay
    from sklearn.cluster import DBSCAN
    g = {"np": np, "scipy": scipy, "Delaunay": Delaunay, "ConvexHull": ConvexHull, "DBSCAN": DBSCAN}
    extract_functions(os.path.join(ref, "voxel_instance_voting.py"), ["in_hull", "min_bounding_box_3d", "cluster"], g)
    # local map: background + K object blobs, predictions 0/1/2
    K = 6
    centers = np.stack([rng.uniform(-30, 30, K), rng.uniform(-30, 30, K), rng.uniform(-1.5, 0.0, K)], -1)
    blobs, blob_pred = [], []
    for k in range(K):
        n = int(rng.integers(150, 400))
        blobs.append(centers[k] + rng.uniform(-1, 1, (n, 3)) * np.array([1.5, 0.8, 0.7]))
        p_dyn = (0.10, 0.60, 0.20, 0.50, 0.30, 0.45)[k]  # 2 votes per dynamic point: threshold is 1/3
        blob_pred.append(np.where(rng.uniform(0, 1, n) < p_dyn, 2, 1))
    bg = np.stack([rng.uniform(-50, 50, 30000), rng.uniform(-50, 50, 30000), rng.uniform(-4, 2, 30000)], -1)
    local_pts = np.concatenate(blobs + [bg]).astype(np.float32)
    local_pts = np.concatenate([local_pts, rng.uniform(0, 1, (len(local_pts), 1)).astype(np.float32)], 1)
    local_pred = np.concatenate(blob_pred + [rng.integers(0, 3, 30000)]).astype(np.int64)
    # AABBs as cluster() builds them: hull corners in float32, z floor lifted by 0.2 (:169-175)
    corners = []
    for k in range(K):
        c = g["min_bounding_box_3d"](blobs[k].astype(np.float32))
        z_min = np.min(c[:, -1])
        c[np.where(c[:, -1] == z_min), -1] += 0.2
        corners.append(c)
    corners = np.stack(corners)
    assert corners.dtype == np.float32
    # put a handful of local-map points EXACTLY on box faces to pin the inclusive comparison
    for k in range(K):
        lo, hi = corners[k].min(0), corners[k].max(0)
        local_pts[30000 + k, :3] = [lo[0], (lo[1] + hi[1]) / 2, (lo[2] + hi[2]) / 2]
        local_pts[30010 + k, :3] = [(lo[0] + hi[0]) / 2, hi[1], (lo[2] + hi[2]) / 2]
        local_pred[30000 + k] = 2
        local_pred[30010 + k] = 1
    stat, dyn, lab = [], [], []
    for k in range(K):
        flag = g["in_hull"](local_pts[:, :3], corners[k])          # voxel_instance_voting.py:177
        pred_in = local_pred[flag]                                  # :180
        s = sum(pred_in[pred_in == 1])                              # :182
        d = sum(pred_in[pred_in == 2])                              # :183
        stat.append(int(s)); dyn.append(int(d)); lab.append(2 if d > s else 1)  # :184-187
    # full cluster() run (DBSCAN + hull + vote + relabel) for an end-to-end check of the vote block
    cur_pts = np.concatenate(blobs + [bg[:5000]]).astype(np.float32)
    cur_pts = np.concatenate([cur_pts, np.zeros((len(cur_pts), 1), np.float32)], 1)
    n_obj = sum(len(b) for b in blobs)
    cur_pred = np.concatenate([np.ones(n_obj, np.int64), rng.integers(0, 3, 5000)])
    cur_bf = np.concatenate([np.full(n_obj, 2, np.uint32), np.zeros(5000, np.uint32)])
    out = g["cluster"](cur_pts.copy(), cur_pred.copy(), cur_bf, local_pts, local_pred)
    np.savez_compressed(os.path.join(GOLD, "instance_a.npz"), local_pts=local_pts, local_pred=local_pred,
                        corners=corners, stat=np.array(stat), dyn=np.array(dyn), label=np.array(lab),
                        cur_pts=cur_pts, cur_pred=cur_pred, cur_bf=cur_bf, cluster_out=out)
    return ["instance_a"]


def gen_stream_vote(ref, rng):
    """One frame of the voxel_voting.py loop body (lines 176-244, `id >= frames_num_max` branch) with the
    reference's own Trans, Crop, Quantize, determine_voxel_labels and get_point_labels_from_voxel_labels."""
    g = {"torch": torch, "np": np}
    extract_functions(os.path.join(ref, "voxel_voting.py"),
                      ["get_point_labels_from_voxel_labels", "determine_voxel_labels", "Quantize"], g)
    gu = {"np": np}
    extract_functions(os.path.join(ref, "datasets", "utils.py"), ["Trans"], gu)
    tr = load_by_path("ref_transforms", os.path.join(ref, "utils", "transforms.py"))
    fov_xyz = [[-50, -50, -4], [50, 50, 2]]                       # voxel_voting.py:138
    crop_to_fov = tr.Crop(dims=(0, 1, 2), fov=fov_xyz)            # :139
    frames_num_max = 8                                            # :140
    size = (64, 64, 12)
    n = 6000
    # synthetic sequence: the ego vehicle drives and turns; points straddle the crop box and voxel borders
    poses, scans, preds = [], [], []
    for t in range(frames_num_max + 1):
        yaw = 0.02 * t
        c, s_ = np.cos(yaw), np.sin(yaw)
        pose = np.array([[c, -s_, 0.0005 * t, 0.9 * t], [s_, c, -0.0003 * t, 0.05 * t * t],
                         [-0.0005 * t, 0.0003 * t, 1.0, 0.01 * t], [0, 0, 0, 1]], dtype=np.float64)
        poses.append(pose)
        pts = np.stack([rng.uniform(-58, 58, n), rng.uniform(-58, 58, n), rng.uniform(-5, 3, n),
                        rng.uniform(0, 1, n)], -1).astype(np.float32)
        pts[: n // 2, :3] *= np.array([0.25, 0.25, 0.6], np.float32)   # dense near the sensor: shared voxels
        pts[:20, 0] = np.float32(50 - 1e-4)                            # on the open crop boundary
        pts[20:40, 1] = np.float32(-50 + 1e-4)
        scans.append(pts)
        preds.append(rng.integers(0, 3, n).astype(np.uint32))
    idx = frames_num_max
    current_points, current_pred_result = scans[idx], preds[idx]
    current_pose_inv = np.linalg.inv(poses[idx])                       # :179
    history_points_list, history_pred_result_list, pose_diffs, transformed = [], [], [], []
    for history_id in np.arange(idx - 1, idx - frames_num_max - 1, -1):   # :182
        pose_diff = current_pose_inv.dot(poses[history_id])            # :187
        hp = gu["Trans"](scans[history_id], pose_diff)                  # :188
        pose_diffs.append(pose_diff)
        transformed.append(hp)
        history_points_list.append(hp)
        history_pred_result_list.append(preds[history_id])
    history_points = np.concatenate(history_points_list, axis=0)       # :193
    history_pred_result = np.concatenate(history_pred_result_list, axis=0)
    current_pred_result_orin = current_pred_result.copy()              # :216
    history_points_t = torch.tensor(history_points)                     # :218 (CPU here)
    history_pred_t = torch.tensor(history_pred_result.astype("uint8"))
    current_points_t = torch.tensor(current_points)
    current_pred_t = torch.tensor(current_pred_result.astype("uint8"))
    history_points_t, history_pred_t, _ = crop_to_fov(history_points_t, history_pred_t)    # :225
    current_points_t, current_pred_t, mask = crop_to_fov(current_points_t, current_pred_t)  # :226
    history_points_num = len(history_points_t)
    local_map_points = torch.cat((history_points_t, current_points_t), dim=0)              # :229
    local_map_prediction = torch.cat((history_pred_t, current_pred_t), dim=0)
    pcds_coord_voxel = g["Quantize"](local_map_points, range_x=(-50.0, 50.0), range_y=(-50.0, 50.0),
                                     range_z=(-4.0, 2.0), size=size)                        # :234
    pcds_coord_cur = pcds_coord_voxel[history_points_num:]
    voxel_label = g["determine_voxel_labels"](pcds_coord_voxel.to(torch.int64), local_map_prediction.to(torch.int64),
                                              size)                                          # :240
    pred_result_new = g["get_point_labels_from_voxel_labels"](pcds_coord_cur.to(torch.int64), voxel_label, size)
    current_pred_result_orin[mask.numpy()] = pred_result_new.numpy()                        # :244
    # history order in the fixture: oldest-independent — stored in the order the script visits them
    hist_ids = list(np.arange(idx - 1, idx - frames_num_max - 1, -1))
    np.savez_compressed(os.path.join(GOLD, "stream_vote_a.npz"),
                        scans=np.stack([scans[i] for i in hist_ids] + [scans[idx]]),
                        preds=np.stack([preds[i] for i in hist_ids] + [preds[idx]]).astype(np.uint8),
                        pose_diffs=np.stack(pose_diffs), transformed=np.stack(transformed),
                        poses=np.stack([poses[i] for i in hist_ids] + [poses[idx]]),
                        size=np.array(size), voxel_labels=voxel_label.numpy().astype(np.uint8),
                        point_labels=current_pred_result_orin.astype(np.int64), n_cropped=int(mask.sum()))
    return ["stream_vote_a"]


def gen_point_stem(ref, rng):
    """networks/backbone.py PointNetStacker(7, 64, pre_bn=True, stack_num=2) exactly as models/StreamMOS.py:77 builds
    it, in eval mode with non-trivial BatchNorm statistics, on loader-shaped 7-channel point features."""
    bb = load_by_path("ref_backbone", os.path.join(ref, "networks", "backbone.py"))
    g = torch.Generator().manual_seed(1234)
    m = bb.PointNetStacker(7, 64, pre_bn=True, stack_num=2)
    bns = [mod for mod in m.modules() if isinstance(mod, torch.nn.BatchNorm2d)]
    convs = [mod for mod in m.modules() if isinstance(mod, torch.nn.Conv2d)]
    assert len(bns) == 3 and len(convs) == 2
    with torch.no_grad():
        for bn in bns:
            c = bn.num_features
            bn.weight.copy_(torch.rand(c, generator=g) * 1.5 + 0.25)
            bn.bias.copy_(torch.randn(c, generator=g) * 0.3)
            bn.running_mean.copy_(torch.randn(c, generator=g) * 0.5)
            bn.running_var.copy_(torch.rand(c, generator=g) * 2.0 + 0.1)
        # the input BatchNorm sees raw metres: statistics of that scale
        bns[0].running_mean.copy_(torch.tensor([0.5, -0.3, -1.2, 0.3, 18.0, 0.5, 0.5]))
        bns[0].running_var.copy_(torch.tensor([300.0, 280.0, 0.8, 0.05, 150.0, 0.08, 0.08]))
    m.eval()
    T, N = 3, 4000
    x = np.stack([rng.uniform(-50, 50, (T, N)), rng.uniform(-50, 50, (T, N)), rng.uniform(-4, 2, (T, N)),
                  rng.uniform(0, 1, (T, N)), rng.uniform(1, 70, (T, N)), rng.uniform(0, 1, (T, N)),
                  rng.uniform(0, 1, (T, N))], 1).astype(np.float32)[..., None]          # (T, 7, N, 1)
    x[:, :3, -50:] = np.array([-1000.0, -1000.0, -4000.0], np.float32)[None, :, None, None]  # loader pads
    with torch.no_grad():
        out = m(torch.from_numpy(x)).numpy()
        out64 = m.double()(torch.from_numpy(x).double()).numpy()
        m.float()
    p = {}
    for i, bn in enumerate(bns):
        p.update({"bn%d_weight" % i: bn.weight.detach().numpy(), "bn%d_bias" % i: bn.bias.detach().numpy(),
                  "bn%d_mean" % i: bn.running_mean.numpy(), "bn%d_var" % i: bn.running_var.numpy(),
                  "bn%d_eps" % i: np.float64(bn.eps)})
    np.savez_compressed(os.path.join(GOLD, "point_stem_a.npz"), x=x, w1=convs[0].weight.detach().numpy(),
                        w2=convs[1].weight.detach().numpy(), out=out, out64=out64,
                        state_keys=np.array(sorted(m.state_dict().keys())), **p)
    return ["point_stem_a"]


def gen_form_batch(ref, rng):
    """The val loader's form_batch (datasets/data_StreamMOS.py:471-493) with the reference's own utils.Quantize,
    utils.SphereQuantize and make_point_feat (AST-extracted: the module imports the dataset stack) on raw, range
    filtered and padded float32 points of T frames, for two TTA sign pairs (form_batch_tta :495-513)."""
    du = load_by_path("ref_dutils", os.path.join(ref, "datasets", "utils.py"))
    g = {"np": np}
    extract_functions(os.path.join(ref, "datasets", "data_StreamMOS.py"), ["make_point_feat"], g)

    class Voxel:  # config/StreamMOS.py:11-19
        RV_theta = (-25.0, 3.0)
        range_x, range_y, range_z = (-50.0, 50.0), (-50.0, 50.0), (-4.0, 2.0)
        bev_shape, rv_shape = (512, 512, 30), (64, 2048)

    T, N, n_valid = 3, 3000, 2700
    r = np.abs(rng.standard_normal((T, N))) * 18.0
    th = rng.uniform(0, 2 * np.pi, (T, N))
    pts = np.stack([np.clip(r * np.cos(th), -49.99, 49.99), np.clip(r * np.sin(th), -49.99, 49.99),
                    np.clip(rng.normal(-1.5, 0.6, (T, N)), -3.99, 1.99), rng.uniform(0, 1, (T, N))], -1).astype(np.float32)
    pts[:, :10, :3] = 0.0                                  # the sensor origin: dist = 1e-12 exactly
    pts[:, 10:20, 0] = np.float32(-50.0)                   # exactly on the lower bound -> x_quan = 0
    pts[:, n_valid:, :] = -1000.0                          # loader pads (data_StreamMOS.py:568-571)
    pts[:, n_valid:, 2] = -4000.0
    out = {}
    for tag, (xs, ys) in {"pp": (1, 1), "mp": (-1, 1)}.items():
        total = pts.reshape(T * N, 4).copy()
        total[:, 0] *= xs
        total[:, 1] *= ys
        xyzi = total[:, :4]
        coord = du.Quantize(xyzi, range_x=Voxel.range_x, range_y=Voxel.range_y, range_z=Voxel.range_z, size=Voxel.bev_shape)
        sphere = du.SphereQuantize(xyzi, phi_range=(-180.0, 180.0), theta_range=Voxel.RV_theta, size=Voxel.rv_shape)
        feat = g["make_point_feat"](xyzi, coord, sphere, Voxel)
        assert feat.dtype == np.float32 and coord.dtype == np.float32
        out["feat_" + tag] = np.ascontiguousarray(feat.reshape(T, N, 7).transpose(0, 2, 1))   # (T, 7, N)
        out["coord_" + tag] = coord.reshape(T, N, 3)
    np.savez_compressed(os.path.join(GOLD, "form_batch_a.npz"), points=pts, **out)
    return ["form_batch_a"]


def gen_sphere(ref, rng):
    """utils.SphereQuantize (datasets/utils.py:172-192) as the val loader's form_batch calls it
    (datasets/data_StreamMOS.py:481-484), run by the reference's own function on float32 points: a 64-beam-like cloud,
    points on the axes (arctan2 at 0, +-pi/2, pi), the sensor origin (dist = 1e-12), the loader's padding rows, and the
    two TTA sign pairs of form_batch_tta. The outputs are HOST-SPECIFIC in their last bit (numpy's float32 arctan2 /
    arcsin): the tests bound the distance instead of asking for equality."""
    du = load_by_path("ref_dutils", os.path.join(ref, "datasets", "utils.py"))
    N, n_valid = 6000, 5600
    r = np.abs(rng.standard_normal(N)) * 18.0 + 0.5
    az = rng.uniform(-np.pi, np.pi, N)
    el = np.deg2rad(rng.uniform(-25.0, 3.0, N))
    pts = np.stack([r * np.cos(el) * np.cos(az), r * np.cos(el) * np.sin(az), r * np.sin(el), rng.uniform(0, 1, N)],
                   -1).astype(np.float32)
    pts[:8, :3] = 0.0
    pts[8:16, :3] = np.array([[0, 1, 0], [0, -1, 0], [1, 0, 0], [-1, 0, 0], [0, -5, -1], [-1e-6, -5, 0.2], [1e-6, -5, 0.2],
                              [3, 3, -1.5]], np.float32)
    pts[n_valid:, :] = -1000.0
    pts[n_valid:, 2] = -4000.0
    out = {}
    for tag, (xs, ys) in {"pp": (1, 1), "mp": (-1, 1), "pm": (1, -1)}.items():
        q = pts.copy()
        q[:, 0] *= xs
        q[:, 1] *= ys
        sph = du.SphereQuantize(q[:, :4], phi_range=(-180.0, 180.0), theta_range=(-25.0, 3.0), size=(64, 2048))
        assert sph.dtype == np.float32, sph.dtype
        out["sphere_" + tag] = sph
    np.savez_compressed(os.path.join(GOLD, "sphere_a.npz"), points=pts, numpy_version=np.array(np.__version__), **out)
    return ["sphere_a"]


def gen_ingest(ref, rng):
    """The val loader's per-frame steps in front of form_batch (datasets/data_StreamMOS.py:515-574) with the
    reference's own utils.Trans and utils.filter_pcds_mask; the compaction and padding lines of __getitem__ (:551-571)
    are restated here because they live inline in the Dataset class. Three raw frames of different lengths, pose_diff =
    inv(pose_cur).dot(pose_ht) as the loader computes it (:427-447; the current frame's own pose_diff is only nearly
    the identity), a fourth frame without any transform and with points exactly on the range bounds."""
    du = load_by_path("ref_dutils", os.path.join(ref, "datasets", "utils.py"))
    rx, ry, rz, n_out = (-50.0, 50.0), (-50.0, 50.0), (-4.0, 2.0), 6144

    def pose(yaw, pitch, tx, ty, tz):
        cy, sy, cp, sp = np.cos(yaw), np.sin(yaw), np.cos(pitch), np.sin(pitch)
        R = np.array([[cy, -sy, 0], [sy, cy, 0], [0, 0, 1]]) @ np.array([[cp, 0, sp], [0, 1, 0], [-sp, 0, cp]])
        P = np.eye(4)
        P[:3, :3], P[:3, 3] = R, (tx, ty, tz)
        return P

    poses = [pose(0.31, 0.004, 120.3, -45.2, 1.1), pose(0.295, 0.003, 119.5, -45.5, 1.08), pose(0.28, 0.002, 118.6, -45.7, 1.07)]
    cur_inv = np.linalg.inv(poses[0])
    raws, diffs = [], []
    for ht, n in enumerate((5000, 5613, 4801)):
        r = np.abs(rng.standard_normal(n)) * 22.0
        th = rng.uniform(0, 2 * np.pi, n)
        raw = np.stack([r * np.cos(th), r * np.sin(th), rng.normal(-1.2, 1.4, n), rng.uniform(0, 1, n)], -1).astype(np.float32)
        raw[::97, 0] = rng.uniform(49.5, 50.5, len(raw[::97]))      # around the upper x bound after alignment
        raw[5::89, 2] = rng.uniform(-4.3, -3.7, len(raw[5::89]))    # around the lower z bound
        raws.append(raw)
        diffs.append(cur_inv.dot(poses[ht]))
    # frame without a pose (the product's "no transform" path): bounds hit exactly
    raw = np.stack([rng.uniform(-55, 55, 3000), rng.uniform(-55, 55, 3000), rng.uniform(-5, 3, 3000), rng.uniform(0, 1, 3000)],
                   -1).astype(np.float32)
    raw[:8, 0] = [-50.0, 50.0, np.nextafter(np.float32(50.0), np.float32(0)), -50.0, 0, 0, 0, 0]
    raw[:8, 1] = [0, 0, 0, 0, -50.0, 50.0, 0, 0]
    raw[:8, 2] = [0, 0, 0, 0, 0, 0, -4.0, 2.0]
    raws.append(raw)
    diffs.append(None)
    outs, counts, masks = [], [], []
    for raw, d in zip(raws, diffs):
        pc = du.Trans(raw, d) if d is not None else raw.copy()                                   # :522
        mask = du.filter_pcds_mask(pc, range_x=rx, range_y=ry, range_z=rz)                       # :545-548
        pc = pc[mask]                                                                            # :551
        pad_length = n_out - pc.shape[0]                                                         # :558
        assert pad_length > 0                                                                    # :559
        pc = np.pad(pc, ((0, pad_length), (0, 0)), 'constant', constant_values=-1000)            # :560
        pc[-pad_length:, 2] = -4000                                                              # :561
        assert pc.dtype == np.float32
        outs.append(pc)
        counts.append(int(mask.sum()))
        masks.append(mask)
    nmax = max(len(r) for r in raws)
    raw_pad = np.zeros((len(raws), nmax, 4), np.float32)
    mask_pad = np.zeros((len(raws), nmax), bool)
    for i, (r, m) in enumerate(zip(raws, masks)):
        raw_pad[i, :len(r)] = r
        mask_pad[i, :len(r)] = m
    np.savez_compressed(os.path.join(GOLD, "ingest_a.npz"), raw=raw_pad, n_raw=np.array([len(r) for r in raws]),
                        pose_diff=np.stack([d if d is not None else np.full((4, 4), np.nan) for d in diffs]),
                        out=np.stack(outs), count=np.array(counts), mask=mask_pad, n_out=np.array(n_out))
    return ["ingest_a"]


def gen_cluster(ref, rng):
    """cluster() of voxel_instance_voting.py:144-193 end to end (sklearn DBSCAN + scipy hull/Delaunay through the
    reference's own code) on a scan whose moving points are interleaved with the rest: objects of several sizes and
    densities (some below the 30-point cut, some touching, one thinner than the 0.2 floor lift), loose noise."""
    import scipy
    from scipy.spatial import ConvexHull, Delaunay
    from sklearn.cluster import DBSCAN
    g = {"np": np, "scipy": scipy, "Delaunay": Delaunay, "ConvexHull": ConvexHull, "DBSCAN": DBSCAN}
    extract_functions(os.path.join(ref, "voxel_instance_voting.py"), ["in_hull", "min_bounding_box_3d", "cluster"], g)
    objs = []
    for k in range(14):
        c = np.array([rng.uniform(-35, 35), rng.uniform(-35, 35), rng.uniform(-1.6, -0.2)])
        n = int(rng.choice([8, 20, 28, 31, 45, 90, 200, 420, 700]))
        ext = np.array([rng.uniform(0.2, 1.2), rng.uniform(0.2, 0.6), rng.uniform(0.15, 0.5)]) * (n / 200.0) ** (1 / 3) * (2.0 if k % 2 else 1.0)
        objs.append(c + rng.uniform(-1, 1, (n, 3)) * ext)
    objs.append(objs[5][:60] + np.array([0.35, 0.0, 0.0]))                      # touches object 5
    objs.append(np.array([12.0, -7.0, -1.0]) + rng.uniform(-1, 1, (150, 3)) * np.array([1.2, 0.6, 0.05]))  # thin
    noise = np.stack([rng.uniform(-40, 40, 900), rng.uniform(-40, 40, 900), rng.uniform(-3, 1, 900)], -1)
    fg = np.concatenate(objs + [noise])
    bg = np.stack([rng.uniform(-50, 50, 16000), rng.uniform(-50, 50, 16000), rng.uniform(-4, 2, 16000)], -1)
    cur = np.concatenate([fg, bg])
    bf = np.concatenate([np.full(len(fg), 2, np.uint32), rng.integers(0, 2, len(bg)).astype(np.uint32)])
    pred = np.concatenate([rng.integers(1, 3, len(fg)), rng.integers(0, 3, len(bg))]).astype(np.int64)
    perm = rng.permutation(len(cur))
    cur, bf, pred = cur[perm].astype(np.float32), bf[perm], pred[perm]
    cur = np.concatenate([cur, rng.uniform(0, 1, (len(cur), 1)).astype(np.float32)], 1)
    # local map = 3 jittered copies of the objects with per-object dynamic rates + background
    lm, lp = [], []
    for k, o in enumerate(objs):
        p_dyn = rng.choice([0.15, 0.3, 0.36, 0.5, 0.7])
        for rep in range(3):
            lm.append(o + rng.normal(0, 0.05, o.shape))
            lp.append(np.where(rng.uniform(0, 1, len(o)) < p_dyn, 2, 1))
    lm.append(np.stack([rng.uniform(-50, 50, 40000), rng.uniform(-50, 50, 40000), rng.uniform(-4, 2, 40000)], -1))
    lp.append(rng.integers(0, 3, 40000))
    local_pts = np.concatenate(lm).astype(np.float32)
    local_pts = np.concatenate([local_pts, np.zeros((len(local_pts), 1), np.float32)], 1)
    local_pred = np.concatenate(lp).astype(np.int64)
    fg_index = np.where(bf == 2)[0]
    fg_labels = DBSCAN(eps=0.3, min_samples=5).fit_predict(cur[fg_index][:, :3])       # :150-153
    out = g["cluster"](cur.copy(), pred.copy(), bf, local_pts, local_pred)
    assert (out != pred).sum() > 100 and len(np.unique(fg_labels)) > 10
    np.savez_compressed(os.path.join(GOLD, "cluster_a.npz"), cur_pts=cur, cur_pred=pred, cur_bf=bf,
                        local_pts=local_pts, local_pred=local_pred, fg_labels=fg_labels.astype(np.int32),
                        cluster_out=out)
    return ["cluster_a"]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", default="/root/reference")
    ap.add_argument("--only", default="", help="generate one family only (e.g. point_stem) and leave the others")
    a = ap.parse_args()
    os.makedirs(GOLD, exist_ok=True)
    torch.manual_seed(0)
    torch.set_num_threads(1)
    rng = np.random.default_rng(20261018)
    made = []
    single = {"point_stem": (gen_point_stem, 99), "form_batch": (gen_form_batch, 55), "cluster": (gen_cluster, 33),
              "msda_module": (gen_msda_module, 0), "ingest": (gen_ingest, 66), "sphere": (gen_sphere, 44)}
    if a.only in single:
        made += single[a.only][0](a.ref, np.random.default_rng(single[a.only][1]))
        for m in made:
            print("%-28s %8.1f KB" % (m, os.path.getsize(os.path.join(GOLD, m + ".npz")) / 1024))
        return
    made += gen_pool(a.ref, rng)
    made += gen_bilinear(a.ref, rng)
    made += gen_msda(a.ref, rng)
    made += gen_voting(a.ref, rng)
    made += gen_instance(a.ref, rng)
    made += gen_stream_vote(a.ref, np.random.default_rng(77))
    made += gen_point_stem(a.ref, np.random.default_rng(99))
    made += gen_form_batch(a.ref, np.random.default_rng(55))
    made += gen_cluster(a.ref, np.random.default_rng(33))
    made += gen_ingest(a.ref, np.random.default_rng(66))
    made += gen_sphere(a.ref, np.random.default_rng(44))
    made += gen_msda_module(a.ref, None)  # last: it re-seeds torch's generator
    for m in made:
        p = os.path.join(GOLD, m + ".npz")
        print("%-28s %8.1f KB" % (m, os.path.getsize(p) / 1024))


if __name__ == "__main__":
    main()
