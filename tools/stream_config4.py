"""BASELINE.json configs[3] / SURVEY 8d "Config 4": a 4071-scan synthetic sequence (the length of SemanticKITTI seq 08)
streamed through the UNMODIFIED reference model `AttNet.infer` (baseline/_ref/StreamMOS, tools/install_ref.py) with
`streammos_b200.dropin.install()`, the short-term memory (`query_embed_store`) carried from scan to scan
(val_StreamMOS.py:85-95), followed per scan by the long-term voting of the produced predictions over an 8-scan window
(voxel_voting.py:176-244 through `voting.quantize_staged` / `determine_voxel_labels` /
`get_point_labels_from_voxel_labels` on a ring resident in HBM).

    python tools/stream_config4.py [--scans 4071] [--points 120000] [--ref-scans 300]

End to end: every scan's loader tensors start in pinned HOST memory (19 MB: 7-channel point features, BEV and range-view
coordinates of T = 3 frames), the voted labels end in pinned host memory; wall clock over the whole run. Hot-path-only
time: CUPTI durations of this repo's kernels (everything the library launches) over a window of scans, against all
device time of the same window (the rest is the reference's convolutions / linears / elementwise torch kernels).
`--ref-scans` > 0 also times the same model on the reference's OWN CUDA extensions (compiled for sm_100a) for
comparison (network only: the reference's voting runs from files on the CPU).
"""
import argparse
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import refmodel  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--scans", type=int, default=4071)
ap.add_argument("--points", type=int, default=120000)
ap.add_argument("--ref-scans", type=int, default=300)
ap.add_argument("--profile-scans", type=int, default=8)
a = ap.parse_args()
dev = torch.device("cuda:0")
N = a.points
NB = 8  # rotating host batches (distinct synthetic scans)

OUR_KERNELS = ("pool_", "gather_", "msda_", "vote_", "point_labels", "instance_vote", "smos_zero", "point_stem", "quantize_kernel",
               "memory_push", "form_batch", "ingest_", "cl_")


def pinned(batch):
    return {k: v.contiguous().pin_memory() for k, v in batch.items()}


host = [pinned(refmodel.make_batch(900 + i, N)) for i in range(NB)]
h2d_bytes = sum(v.numel() * v.element_size() for v in host[0].values())


def run(mode, scans, vote):
    from streammos_b200 import synthetic, voting
    net, d = refmodel.load_attnet(mode, seed=0)
    devb = [{k: torch.empty_like(v, device=d) for k, v in host[0].items()} for _ in range(2)]
    ring_pts = torch.full((9, N, 4), -1000.0, device=d)
    ring_pred = torch.zeros((9, N), dtype=torch.uint8, device=d)
    out_host = torch.empty(N, dtype=torch.int64).pin_memory()
    copy_s = torch.cuda.Stream(d)
    ready = [torch.cuda.Event() for _ in range(2)]
    freed = [torch.cuda.Event() for _ in range(2)]
    main = torch.cuda.current_stream(d)

    def upload(i):
        j = i % 2
        with torch.cuda.stream(copy_s):
            copy_s.wait_event(freed[j])
            for k, v in host[i % NB].items():
                devb[j][k].copy_(v, non_blocking=True)
            ready[j].record(copy_s)

    def step(i, store):
        j = i % 2
        main.wait_event(ready[j])
        b = devb[j]
        pred_cls, _, _, _, store = net.infer(b, i, store)
        labels = None
        if vote:
            pred = pred_cls[0, :, :, 0].argmax(0).to(torch.uint8)                      # (N,) the network's prediction
            xyzi = b["pcds_xyzi"][0, 0, 0, :4, :, 0].t().contiguous()                  # raw points of the current frame
            q, coords, lab = voting.quantize_staged(ring_pts, ring_pred, synthetic.RANGE_X, synthetic.RANGE_Y,
                                                    synthetic.RANGE_Z, synthetic.BEV_SHAPE, new_points=xyzi, new_pred=pred,
                                                    cur_slot=8, hist_slot=(i - 1) % 8, want_q=False, crop_eps=1e-4)
            vl = voting.determine_voxel_labels(coords, lab, synthetic.BEV_SHAPE, num_classes=3)
            labels = voting.get_point_labels_from_voxel_labels(coords[8 * N:], vl, synthetic.BEV_SHAPE)
            out_host.copy_(labels, non_blocking=True)
        freed[j].record(main)
        return store

    with torch.no_grad():
        store = None
        upload(0)
        for i in range(4):  # warm-up (cuDNN autotune, plan cache learning)
            upload(i + 1)
            store = step(i, store)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for i in range(4, 4 + scans):
            upload(i + 1)
            store = step(i, store)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        share = None
        if mode == "b200" and a.profile_scans > 0:
            from torch.profiler import ProfilerActivity, profile
            with profile(activities=[ProfilerActivity.CUDA]) as prof:
                for i in range(4 + scans, 4 + scans + a.profile_scans):
                    upload(i + 1)
                    store = step(i, store)
                torch.cuda.synchronize()
            evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
            ours = sum(e.time_range.end - e.time_range.start for e in evs if any(t in e.name for t in OUR_KERNELS))
            allk = sum(e.time_range.end - e.time_range.start for e in evs if "memcpy" not in e.name.lower())
            share = (ours / a.profile_scans, allk / a.profile_scans)
    refmodel.purge()
    return dt, share, int(out_host.sum()) if vote else None


dt, share, chk = run("b200", a.scans, vote=True)
print("config 4: %d scans of %d points through the unmodified reference AttNet.infer + long-term voting (8-scan window), 1 x %s"
      % (a.scans, N, torch.cuda.get_device_name(0)))
print("  drop-in (streammos_b200.dropin.install()):   %8.2f s   %7.2f ms/scan   %7.1f scans/s end to end "
      "(H2D %.1f MB + D2H %.2f MB per scan inside)" % (dt, dt / a.scans * 1e3, a.scans / dt, h2d_bytes / 1e6, N * 8 / 1e6))
if share:
    print("  device time per scan (CUPTI, %d scans): %.0f us in this repo's kernels (hot path + stem) of %.0f us in all kernels "
          "(%.1f %%): the rest is the reference's convolutions, linears and elementwise torch kernels"
          % (a.profile_scans, share[0], share[1], 100.0 * share[0] / share[1]))
print("  checksum of the last voted labels: %d" % chk)
if a.ref_scans > 0 and refmodel.ref_ext("msda") and refmodel.ref_ext("point_deep_cuda"):
    dt_r, _, _ = run("cuda_reference", a.ref_scans, vote=False)
    dt_b, _, _ = run("b200", a.ref_scans, vote=False)
    print("  network only, %d scans: reference CUDA extensions (sm_100a build) %7.2f ms/scan   drop-in %7.2f ms/scan"
          % (a.ref_scans, dt_r / a.ref_scans * 1e3, dt_b / a.ref_scans * 1e3))
