"""Run the fused raw-scan PointNet stem a few times (for ncu captures of point_stem_umma_kernel).
    python tools/stem_only.py [--iters 5]"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from streammos_b200 import ops, stream, synthetic  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--iters", type=int, default=5)
ap.add_argument("--points", type=int, default=120000)
a = ap.parse_args()
dev = torch.device("cuda:0")
hot = stream.HotPath(dev, a.points, seed=0)
raw = [stream.make_host_raw_scan(i, a.points, pin=False).to(dev) for i in range(2)]
args = (synthetic.RANGE_X, synthetic.RANGE_Y, synthetic.RANGE_Z, hot.size) + tuple(hot.stem.fused_parameters())
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
with torch.no_grad():
    for i in range(3):
        ops.point_stem_forward_raw(raw[i % 2].points, *args, point_major_out=True)
    torch.cuda.synchronize()
    e0.record()
    for i in range(a.iters):
        y, c = ops.point_stem_forward_raw(raw[i % 2].points, *args, point_major_out=True)
    e1.record()
torch.cuda.synchronize()
print("stem (raw, point-major out): %.1f us per call, checksum %.3f" % (e0.elapsed_time(e1) * 1e3 / a.iters, float(y.sum())))
