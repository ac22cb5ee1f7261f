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
ap.add_argument("--channel-major-feat", action="store_true",
                help="hand pool #1 the (B, C, N, 1)-contiguous features of the reference stem instead of point-major ones")
ap.add_argument("--keep-plans", action="store_true", help="let the plan cache serve the repeated scans (no plan kernels)")
a = ap.parse_args()
dev = torch.device("cuda:0")
hot = stream.HotPath(dev, a.points, seed=0, branches=False, vote_api=a.vote_api)
if a.loader:
    scans = [stream.make_host_loader_scan(i, a.points).to(dev) for i in range(4)]
else:
    scans = [stream.make_host_scan(i, a.points, feat_point_major=not a.channel_major_feat).to(dev) for i in range(4)]
from streammos_b200 import plan_cache  # noqa: E402


def run(i):
    if not a.keep_plans:  # forget the plans (not what was learned): every step rebuilds them in its three batches
        plan_cache._entries.clear()
    fn(scans[i % 4])

fn = {"step": hot.step, "vote": hot.long_term_voting, "proj": hot.projection}[a.what]
with torch.no_grad():
    for i in range(8):  # the plan cache learns its prefetch batches on the first scans: profile the steady state
        run(i)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        for i in range(a.steps):
            run(i)
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
