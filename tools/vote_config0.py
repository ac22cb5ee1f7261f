"""BASELINE.json configs[0]: the reference's own CPU-runnable case — voxel_voting.py's long-term voting of one scan
against the 8 scans before it, ~120 k points each — run by the REFERENCE'S OWN FUNCTIONS on the host cores, next to the
same frames through `voting.StreamingVoter` (smos_vote_stream) on the B200, labels compared bit for bit.

    python tools/vote_config0.py [--frames 12] [--ref baseline/_ref/StreamMOS]

Reference side = the loop body of voxel_voting.py:176-244 (`id >= frames_num_max` branch) with the functions extracted
from the reference tree that travelled to this box (tools/install_ref.py): `utils.Trans` (datasets/utils.py:116-126),
`transforms.Crop` (utils/transforms.py:151-161), `Quantize`, `determine_voxel_labels`,
`get_point_labels_from_voxel_labels` (voxel_voting.py:38-91), on torch CPU tensors with all host threads (the script
itself moves the tensors to a GPU; without one this is what it computes). File I/O is left out on both sides: the scans
and predictions start in host memory. Device side: the new raw scan + predictions cross PCIe from pinned memory, the 8
older scans are resident in HBM, the refined labels return to pinned memory — copies inside the timed region.
"""
import argparse
import ast
import importlib.util
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

ap = argparse.ArgumentParser()
ap.add_argument("--frames", type=int, default=12, help="frames voted (after the 8 warm-up frames)")
ap.add_argument("--ref", default=os.path.join(ROOT, "baseline", "_ref", "StreamMOS"))
a = ap.parse_args()

from streammos_b200 import synthetic, voting  # noqa: E402


def extract(path, names, glob):
    tree = ast.parse(open(path).read())
    body = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in names]
    assert {n.name for n in body} == set(names), path
    exec(compile(ast.Module(body=body, type_ignores=[]), path, "exec"), glob)
    return glob


def load_by_path(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


g = extract(os.path.join(a.ref, "voxel_voting.py"),
            ["get_point_labels_from_voxel_labels", "determine_voxel_labels", "Quantize"], {"torch": torch, "np": np})
gu = extract(os.path.join(a.ref, "datasets", "utils.py"), ["Trans"], {"np": np})
tr = load_by_path("ref_transforms", os.path.join(a.ref, "utils", "transforms.py"))
crop_to_fov = tr.Crop(dims=(0, 1, 2), fov=[[-50, -50, -4], [50, 50, 2]])   # voxel_voting.py:138-139
FRAMES_MAX, SIZE = 8, (512, 512, 30)                                        # :140, :233

# ---- a synthetic drive: raw scans in their own sensor frame, a pose per scan, a prediction per point ------------------
n_total = FRAMES_MAX + a.frames
rng = np.random.default_rng(2026)
scans, preds, poses = [], [], []
pose = np.eye(4)
for i in range(n_total):
    scans.append(synthetic.lidar_raw_scan(np.random.default_rng(500 + i)))
    preds.append(rng.integers(0, 3, len(scans[-1])).astype(np.uint32))
    yaw = 0.004
    step = np.eye(4)
    step[:2, :2] = [[np.cos(yaw), -np.sin(yaw)], [np.sin(yaw), np.cos(yaw)]]
    step[:3, 3] = (0.62, 0.03, 0.001)
    pose = pose.dot(step)
    poses.append(pose.copy())


def reference_frame(idx):
    """voxel_voting.py:177-244 for frame `idx` (>= FRAMES_MAX), tensors on the CPU."""
    current_points, current_pred_result = scans[idx], preds[idx]
    current_pose_inv = np.linalg.inv(poses[idx])
    hp, hl = [], []
    for history_id in np.arange(idx - 1, idx - FRAMES_MAX - 1, -1):
        hp.append(gu["Trans"](scans[history_id], current_pose_inv.dot(poses[history_id])))
        hl.append(preds[history_id])
    history_points = torch.tensor(np.concatenate(hp, axis=0))
    history_pred = torch.tensor(np.concatenate(hl, axis=0).astype("uint8"))
    out = current_pred_result.copy()
    cur_t, cur_l = torch.tensor(current_points), torch.tensor(current_pred_result.astype("uint8"))
    history_points, history_pred, _ = crop_to_fov(history_points, history_pred)
    cur_t, cur_l, mask = crop_to_fov(cur_t, cur_l)
    n_hist = len(history_points)
    local_map_points = torch.cat((history_points, cur_t), dim=0)
    local_map_prediction = torch.cat((history_pred, cur_l), dim=0)
    q = g["Quantize"](local_map_points, range_x=(-50.0, 50.0), range_y=(-50.0, 50.0), range_z=(-4.0, 2.0), size=SIZE)
    vl = g["determine_voxel_labels"](q.to(torch.int64), local_map_prediction.to(torch.int64), SIZE)
    new = g["get_point_labels_from_voxel_labels"](q[n_hist:].to(torch.int64), vl, SIZE)
    out[mask.numpy()] = new.numpy()
    return out.astype(np.int64)


cores = torch.get_num_threads()
reference_frame(FRAMES_MAX)  # warm-up (allocator, thread pool)
t0 = time.perf_counter()
ref_out = [reference_frame(i) for i in range(FRAMES_MAX, n_total)]
ref_s = (time.perf_counter() - t0) / a.frames

# ---- the same frames through the resident ring on the device ------------------------------------------------------------
dev = torch.device("cuda:0")
n_cap = (max(len(s) for s in scans) + 255) // 256 * 256
h_pts = [torch.from_numpy(s).pin_memory() for s in scans]
h_pred = [torch.from_numpy(p.astype(np.uint8)).pin_memory() for p in preds]
h_out = [torch.empty(len(s), dtype=torch.int64).pin_memory() for s in scans]


def device_run(timed):
    voter = voting.StreamingVoter(frames_num_max=FRAMES_MAX, size=SIZE)
    for i in range(FRAMES_MAX):  # the window before the first voted frame: resident, not timed
        voter.push(h_pts[i].to(dev), h_pred[i].to(dev), poses[i])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for i in range(FRAMES_MAX, n_total):
        voter.push(h_pts[i].to(dev, non_blocking=True), h_pred[i].to(dev, non_blocking=True), poses[i])
        _, labels = voter.vote()
        h_out[i].copy_(labels, non_blocking=True)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e-3 / a.frames, (time.perf_counter() - t0) / a.frames


device_run(False)
dev_s, wall_s = device_run(True)
same = all(np.array_equal(h_out[i].numpy(), ref_out[i - FRAMES_MAX]) for i in range(FRAMES_MAX, n_total))
changed = float(np.mean([np.mean(ref_out[i - FRAMES_MAX] != preds[i]) for i in range(FRAMES_MAX, n_total)]))
res = {"config": "BASELINE.json configs[0]: long-term voxel voting of one scan against the 8 before it",
       "frames": a.frames, "points_per_scan": int(np.mean([len(s) for s in scans])),
       "reference_cpu": {"ms_per_frame": ref_s * 1e3, "frames_per_s": 1.0 / ref_s, "cores": cores,
                         "code": "voxel_voting.py:176-244 body, the reference's own functions on torch CPU tensors"},
       "b200": {"ms_per_frame_device": dev_s * 1e3, "ms_per_frame_wall": wall_s * 1e3, "frames_per_s": 1.0 / max(dev_s, wall_s),
                "h2d_bytes_per_frame": int(np.mean([len(s) for s in scans])) * 17, "d2h_bytes_per_frame": int(np.mean([len(s) for s in scans])) * 8,
                "code": "voting.StreamingVoter: push + vote (smos_vote_stream), 8 older scans resident in HBM"},
       "labels_bit_identical": bool(same), "fraction_of_labels_changed_by_the_vote": changed}
print(json.dumps(res, indent=1))
assert same, "device labels differ from the reference's"
