"""In-situ kernel durations of the eager serial step (CUPTI through torch.profiler, no replay, warm caches as in a
real stream) — complements the ncu launch list, whose per-launch times are cold-cache and serialised.
    python tools/profile_step.py [--steps 8] [--what step|vote|proj]"""
import argparse
import os
import sys
from collections import OrderedDict

import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from streammos_b200 import stream  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=8)
ap.add_argument("--points", type=int, default=120000)
ap.add_argument("--what", default="step")
ap.add_argument("--vote-api", default="reference")
ap.add_argument("--loader", action="store_true", help="start every step from the loader tensors (PointNet stem on device)")
a = ap.parse_args()
dev = torch.device("cuda:0")
hot = stream.HotPath(dev, a.points, seed=0, branches=False, vote_api=a.vote_api)
scans = [(stream.make_host_loader_scan if a.loader else stream.make_host_scan)(i, a.points).to(dev) for i in range(4)]
fn = {"step": hot.step, "vote": hot.long_term_voting, "proj": hot.projection}[a.what]
with torch.no_grad():
    for i in range(8):  # the plan cache learns its prefetch batches on the first scans: profile the steady state
        fn(scans[i % 4])
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        for i in range(a.steps):
            fn(scans[i % 4])
        torch.cuda.synchronize()
evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
evs.sort(key=lambda e: e.time_range.start)
per = len(evs) // a.steps
agg = OrderedDict()
for k, e in enumerate(evs):
    key = (k % per, e.name[:90])
    agg.setdefault(key, []).append(e.time_range.end - e.time_range.start)
tot = 0.0
for (k, name), v in agg.items():
    v = sorted(v)
    med = v[len(v) // 2]
    tot += med
    print("%3d %-90s %8.1f us" % (k, name, med))
print("device events/step %d, sum of medians %.1f us" % (per, tot))
