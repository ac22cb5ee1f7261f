"""Kernel-by-kernel durations (CUPTI, through torch.profiler) of one voting.cluster() call on a synthetic scan.

    python tools/profile_cluster.py [--objects 30 --per-object 100]
"""
import argparse
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from streammos_b200 import voting  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--objects", type=int, default=30)
ap.add_argument("--per-object", type=int, default=100)
ap.add_argument("--points", type=int, default=120000)
a = ap.parse_args()
dev = torch.device("cuda:0")
rng = np.random.default_rng(0)
N = a.points
objs = [np.array([rng.uniform(-40, 40), rng.uniform(-40, 40), rng.uniform(-1.5, -0.5)]) +
        rng.uniform(-1, 1, (a.per_object, 3)) * np.array([1.0, 0.5, 0.4]) for _ in range(a.objects)]
fgp = np.concatenate(objs + [np.stack([rng.uniform(-50, 50, 300), rng.uniform(-50, 50, 300), rng.uniform(-3, 1, 300)], -1)])
M = len(fgp)
cur = np.concatenate([np.stack([rng.uniform(-50, 50, N), rng.uniform(-50, 50, N), rng.uniform(-4, 2, N)], -1),
                      np.zeros((N, 1))], 1).astype(np.float32)
idx = rng.permutation(N)[:M]
cur[idx, :3] = fgp
bf = np.zeros(N, np.int32)
bf[idx] = 2
local = np.concatenate([np.stack([rng.uniform(-50, 50, 9 * N), rng.uniform(-50, 50, 9 * N), rng.uniform(-4, 2, 9 * N)], -1),
                        np.zeros((9 * N, 1))], 1).astype(np.float32)
cur_t, bf_t, local_t = (torch.from_numpy(x).to(dev) for x in (cur, bf, local))
lpred = torch.randint(0, 3, (9 * N,), device=dev)
pred = torch.randint(0, 3, (N,), device=dev)
for _ in range(3):
    voting.cluster(cur_t, pred.clone(), bf_t, local_t, lpred)
torch.cuda.synchronize()
with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA]) as prof:
    for _ in range(5):
        voting.cluster(cur_t, pred.clone(), bf_t, local_t, lpred)
    torch.cuda.synchronize()
rows = {}
for ev in prof.events():
    if ev.device_type == torch.autograd.DeviceType.CUDA:
        rows.setdefault(ev.name, []).append(ev.device_time)
print("M = %d moving points of %d" % (M, N))
tot = 0.0
for name, ts in rows.items():
    med = sorted(ts)[len(ts) // 2]
    tot += med * len(ts) / 5
    print("%-90s x%d  %8.1f us" % (name[:90], len(ts) // 5, med))
print("sum per call: %.1f us" % tot)
