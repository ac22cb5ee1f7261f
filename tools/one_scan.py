"""Run a few eager hot-path steps (for ncu launch lists / single-kernel captures).
    python tools/one_scan.py [--steps 3] [--channel-major] [--vote-api fused]"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from streammos_b200 import stream  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=3)
ap.add_argument("--points", type=int, default=120000)
ap.add_argument("--channel-major", action="store_true")
ap.add_argument("--vote-api", default="reference")
ap.add_argument("--grids-channels-last", action="store_true")
ap.add_argument("--channel-major-feat", action="store_true",
                help="hand pool #1 the (B, C, N, 1)-contiguous features of the reference stem instead of point-major ones")
a = ap.parse_args()
dev = torch.device("cuda:0")
hot = stream.HotPath(dev, a.points, seed=0, point_major=not a.channel_major, vote_api=a.vote_api,
                     grids_channels_last=a.grids_channels_last)
# one scan per step, so that every step builds its pooling plans (from the second step on in the learned batches)
scans = [stream.make_host_scan(i, a.points, feat_point_major=not a.channel_major_feat).to(dev) for i in range(a.steps)]
torch.cuda.synchronize()
with torch.no_grad():
    for i in range(a.steps):
        labels, sums, _ = hot.step(scans[i])
torch.cuda.synchronize()
print("ok", int(labels.sum()), int(sums.sum()))
