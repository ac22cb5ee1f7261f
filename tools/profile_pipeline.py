"""Timeline of the pipelined graph replays (CUPTI through torch.profiler): how much of a scan's wall time has
0 / 1 / 2+ kernels running, and which kernels the time goes to when they overlap.
    python tools/profile_pipeline.py [--scans 32]
The end-to-end loop of bench.py (copies included) goes through the same analysis:
    SMOS_E2E_PROFILE=out.txt python bench.py --e2e-only"""
import argparse
import os
import sys
from collections import defaultdict

import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from streammos_b200 import pipeline, stream  # noqa: E402

def analyse(prof, n_scans, out=sys.stdout):
    """Concurrency statistics of the CUDA activities (kernels and copies) a torch profiler run recorded."""
    def print(*args):  # noqa: A001 — everything below goes to `out`
        out.write(" ".join(str(x) for x in args) + "\n")
    evs = [(e.time_range.start, e.time_range.end, e.name) for e in prof.events()
           if e.device_type == torch.autograd.DeviceType.CUDA]
    evs.sort()
    t0, t1 = evs[0][0], max(e[1] for e in evs)
    span = t1 - t0
    print("scans %d  span %.1f us  -> %.1f us/scan ; kernels/scan %.1f ; sum of kernel time %.1f us/scan"
          % (n_scans, span, span / n_scans, len(evs) / n_scans, sum(e[1] - e[0] for e in evs) / n_scans))
    # sweep line: time with k kernels active
    pts = []
    for s, e, n in evs:
        pts.append((s, 1))
        pts.append((e, -1))
    pts.sort()
    hist = defaultdict(float)
    k, last = 0, pts[0][0]
    for t, d in pts:
        hist[min(k, 4)] += t - last
        last = t
        k += d
    for kk in sorted(hist):
        print("  %s kernels active: %6.1f us/scan (%4.1f %%)" % (("%d" % kk) if kk < 4 else "4+", hist[kk] / n_scans, 100 * hist[kk] / span))
    # which kernels run ALONE (time with exactly one kernel active, attributed to that kernel)
    active, alone = {}, defaultdict(float)
    ev2 = []
    for idx, (s, e, n) in enumerate(evs):
        ev2.append((s, 0, idx))
        ev2.append((e, 1, idx))
    ev2.sort()
    last = ev2[0][0]
    for tt, kind, idx in ev2:
        if len(active) == 1:
            alone[evs[next(iter(active))][2][:70]] += tt - last
        last = tt
        if kind == 0:
            active[idx] = True
        else:
            active.pop(idx, None)
    print("time with exactly ONE kernel active, by kernel (us/scan):")
    for n, v in sorted(alone.items(), key=lambda x: -x[1])[:14]:
        print("  %-70s %7.1f us" % (n, v / n_scans))
    by = defaultdict(float)
    for s, e, n in evs:
        by[n[:70]] += e - s
    print("in-pipeline kernel time per scan (durations stretch when kernels share the GPU):")
    for n, v in sorted(by.items(), key=lambda x: -x[1])[:16]:
        print("  %-70s %7.1f us" % (n, v / n_scans))


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--scans", type=int, default=32)
    ap.add_argument("--points", type=int, default=120000)
    ap.add_argument("--in-flight", type=int, default=4)
    ap.add_argument("--branches", action="store_true")
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    hot = stream.HotPath(dev, a.points, seed=0, branches=a.branches)
    scans = [stream.make_host_scan(i, a.points).to(dev) for i in range(8)]
    pipe = pipeline.ScanPipeline(hot, scans, use_graphs=True, scans_in_flight=a.in_flight)
    for _ in range(16):
        pipe.submit()
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(a.scans):
            pipe.submit()
        torch.cuda.synchronize()
    analyse(prof, a.scans)
