#!/usr/bin/env python
"""Benchmark of the StreamMOS hot path on B200 (BASELINE.json metric: scans/s at ~120k points, % of HBM roofline).

    python bench.py [--gpus N] [--steps K] [--warmup W]            # this repo's sm_100a kernels
    python bench.py --impl reference [...]                          # torch-CPU port of the reference ops (oracle/cpu_path.py)
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W                      # one rank per GPU, independent scan streams

One "step" = one synthetic ~120k-point scan through the whole hot path: 5 x VoxelMaxPool + 5 x BilinearSample
(cascade projection), 2 x MSDeformAttn forward (temporal fusion), voxel voting over 8+1 scans and per-instance
votes (long-term memory). Rank 0 prints ONE JSON line.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "scans/s at ~120k pts (whole hot path: projection + deformable attention + voting)"
N_SCANS = 8  # distinct synthetic scans cycled through (inputs >> L2: ~100 MB each, ~700 MB touched per step)


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--warmup", type=int, default=50)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--points", type=int, default=120000)
    ap.add_argument("--no-graph", action="store_true", help="time eager launches instead of CUDA-graph replays")
    ap.add_argument("--channel-major", action="store_true",
                    help="keep gathered point features (B,C,N,1)-contiguous instead of point-major")
    ap.add_argument("--vote-api", default="reference", choices=["reference", "fused"])
    ap.add_argument("--cpu-scans", type=int, default=20, help="scans timed for the cpu_baseline leg")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--grids-channels-last", action="store_true",
                    help="CNN stand-in feature maps in channels_last (what a channels_last model hands to the gathers)")
    ap.add_argument("--no-variants", action="store_true", help="skip the extra layout/API variant measurement")
    ap.add_argument("--in-flight", type=int, default=4,
                    help="scans in flight per stream (projection streams); measured on B200: 1 -> 3424, 2 -> 4404, "
                         "4 -> 4680, 8 -> 4535 scans/s")
    ap.add_argument("--breakdown", action="store_true", help="also print a per-operator table to stderr")
    ap.add_argument("--branches", action="store_true",
                    help="experiment: run the four independent operator groups of the projection (independent only because "
                         "the CNN outputs are resident stand-ins) as parallel graph branches; the default is the "
                         "reference's serial chain")
    ap.add_argument("--no-branches", action="store_true", help="accepted for compatibility; the serial chain is the default")
    ap.add_argument("--explicit-plans", action="store_true",
                    help="call the operators with the explicit plan API (plan= / order=, five plans per batch of launches) "
                         "instead of the reference's signatures + plan cache")
    ap.add_argument("--no-ordered-gathers", action="store_true",
                    help="BEV gathers visit the points in scan order instead of the pooling plan's cell order")
    ap.add_argument("--ordered-rv", action="store_true", help="range-view gathers in cell order too")
    ap.add_argument("--gather-taps", action="store_true",
                    help="gathers read their sampling state from records the plan build emits (measured: no gain)")
    ap.add_argument("--e2e-only", action="store_true", help="tuning aid: measure and print only the raw-scan e2e leg")
    ap.add_argument("--families-only", action="store_true",
                    help="tuning aid: print only the per-family device times (not the contract's JSON line) and exit")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            pass
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def dist_setup():
    import torch
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        backend = "nccl" if torch.cuda.is_available() else "gloo"
        kw = {}
        if backend == "nccl":
            torch.cuda.set_device(local)
            kw["device_id"] = torch.device("cuda", local)
        # NCCL prints its version banner (the image sets NCCL_DEBUG=VERSION) with printf to stdout when the first
        # communicator comes up: send fd 1 to stderr while that happens, so stdout carries the ONE JSON line only
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group(backend, rank=rank, world_size=world, **kw)
            dist.barrier()
            if backend == "nccl":
                torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)
    return world, rank, local


def barrier(world):
    if world > 1:
        import torch.distributed as dist
        dist.barrier()


def max_over_ranks(x, world, device):
    """Timing plumbing only (no data-path collective): max of a scalar over ranks."""
    if world == 1:
        return x
    import torch
    import torch.distributed as dist
    t = torch.tensor([x], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


# ------------------------------------------------------------------------------------------------------------
def cpu_state(hot):
    return {"x0": hot.x0.cpu(), "x1": hot.x1.cpu(), "dec": hot.dec.cpu(), "memory": hot.memory.cpu().clone(),
            "local_pts": hot.local_pts.cpu().clone(), "local_pred": hot.local_pred.cpu().clone(),
            "box_lo": hot.box_lo.cpu(), "box_hi": hot.box_hi.cpu(), "scan_index": 0}


def time_cpu_path(state, scans, n_scans, warmup=1, budget_s=120.0):
    """A torch-CPU port of the reference operators (oracle/cpu_path.py: scatter_reduce pooling, F.grid_sample,
    ms_deform_attn_core_pytorch, torch voting) on the host cores: scans/s over a bounded sample (at most `n_scans` scans
    and at most `budget_s` seconds). A reported baseline, not a target."""
    import torch
    from oracle.cpu_path import CpuHotPath
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    path = CpuHotPath(state)
    done = 0
    with torch.no_grad():
        for i in range(warmup):
            path.step(scans[i % len(scans)])
        t0 = time.perf_counter()
        while done < n_scans:
            path.step(scans[(warmup + done) % len(scans)])
            done += 1
            if time.perf_counter() - t0 > budget_s:
                break
        dt = time.perf_counter() - t0
    return done / dt, dt / done * 1e3, cores, done


def run_reference(args, world, rank):
    """`--impl reference`: the reference's own CPU implementation of the path. Rank 0 alone runs it."""
    if rank != 0:
        return
    import torch
    from streammos_b200 import stream
    hot = stream.HotPath("cpu", n_points=args.points, seed=0)
    scans = [stream.make_host_scan(i, args.points, pin=False) for i in range(min(N_SCANS, 4))]
    sps, ms, cores, done = time_cpu_path(cpu_state(hot), scans, args.steps, warmup=max(1, min(args.warmup, 2)))
    sample = ("%d scans timed (1 scan per step; capped at 120 s of the %d requested), oracle/cpu_path.py: torch %s CPU "
              "ops on %d threads" % (done, args.steps, torch.__version__, cores))
    line = {"impl": "reference", "metric": METRIC, "value": sps, "unit": "scans/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args, graph=False, world=1),
            "cpu_baseline": {"value": sps, "unit": "scans/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": sps, "unit": "scans/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def workload_config(args, graph, world):
    return {"workload": "BASELINE.json configs[1]+[2]+[0] per scan (5 VoxelMaxPool + 5 BilinearSample + 2 MSDeformAttn fwd "
                        "+ voxel voting over 8+1 scans + 32 instance votes), B=1, T=3",
            "points_per_scan": args.points, "bev_shape": [512, 512, 30], "rv_shape": [64, 2048],
            "memory": [64, 64, 128], "streams": world, "parallelism": "1 independent scan stream per GPU, no collectives",
            "launch": "cuda-graph replay" if graph else "eager", "vote_api": args.vote_api,
            "point_feature_layout": "channel-major (reference layout: VoxelMaxPool #1 permutes)" if args.channel_major else
                                    "point-major (channels_last strides), stem output included",
            "cnn_grid_layout": "channels_last" if args.grids_channels_last else "NCHW (reference default)",
            "pipeline": "3 graphs per scan (projection / temporal fusion / voting), %d scans in flight on %d CUDA streams; "
                        "cross-scan dependencies (short-term memory, voting ring) enforced with events; %s"
                        % (args.in_flight, args.in_flight + 1,
                           "operators of the network in the reference's serial order; the instance vote "
                           "(voxel_instance_voting.py, an independent post-processing pass of the reference) runs beside the "
                           "voxel vote (voxel_voting.py) as a second graph branch" if not args.branches else
                           "independent operators of a scan are parallel graph "
                           "branches (pool #1 | half-scale chain | quarter-scale chain | gather #5; voxel | instance votes)"),
            "operator_api": "explicit plan API (plan= / order=)" if args.explicit_plans else
                            "reference signatures only (VoxelMaxPool(feat, ind, size, scale), BilinearSample(grid, coord), "
                            "MSDA.ms_deform_attn_forward, Quantize / determine_voxel_labels / get_point_labels_from_voxel_labels); "
                            "plans shared through the plan cache as under the unmodified reference model",
            "gather_order": "scan order" if args.no_ordered_gathers else
                            ("cell order of the shared pooling plan (BEV%s)%s" % (
                                " + RV" if (args.ordered_rv or args.gather_taps) else "",
                                "; sampling state emitted by the plan build" if args.gather_taps else "")),
            "l2": "no explicit flush: %d distinct scans cycled, ~100 MB inputs and ~700 MB touched per step (>126 MB L2)" % N_SCANS}


# ------------------------------------------------------------------------------------------------------------
def run_b200(args, world, rank, local):
    import torch
    from streammos_b200 import _lib, ops, pipeline, stream
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py: no CUDA device — the b200 arm has no CPU fallback (use --impl reference)")
    _lib.load()
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    # one process per GPU on one host: run next to the GPU (cores of its NUMA node) before any pinned buffer is
    # allocated, so submits and H2D reads do not cross the socket link (SMOS_NO_PIN=1: leave the affinity alone)
    from streammos_b200 import multi
    pinned_cpus = None if os.environ.get("SMOS_NO_PIN") else multi.pin_rank_to_gpu(local, world)
    use_graph = not args.no_graph
    hot = stream.HotPath(dev, n_points=args.points, seed=rank, point_major=not args.channel_major,
                         vote_api=args.vote_api, overlap_voting=False,
                         grids_channels_last=args.grids_channels_last, branches=args.branches, batch_plans=args.explicit_plans,
                         ordered_gathers=not args.no_ordered_gathers, ordered_rv=args.ordered_rv,
                         gather_taps=args.gather_taps)
    # the 64-channel point features are what the PointNet stem hands to VoxelMaxPool #1; the drop-in stem
    # (backbone.PointNetStacker, tcgen05 kernel) emits them point-major (channels_last strides), like every other
    # point-feature tensor of the path. --channel-major keeps the reference's (B, C, N, 1)-contiguous layout.
    host = [stream.make_host_scan(rank * 1000 + i, args.points, feat_point_major=not args.channel_major)
            for i in range(N_SCANS)]
    devb = [h.to(dev) for h in host]
    assert devb[0].feat.stride() == host[0].feat.stride()
    torch.cuda.synchronize()
    cpu_hot_state = cpu_state(hot) if (rank == 0 and world == 1 and not args.no_cpu_baseline) else None

    if args.families_only:
        with torch.no_grad():
            hot.step(devb[0])
            hot.scan_index = 0
            raw = None
            if os.environ.get("SMOS_FAMILIES_RAW"):
                raw = [stream.make_host_raw_scan(rank * 1000 + i, args.points, pin=False).to(dev) for i in range(N_SCANS)]
            fam = family_breakdown(hot, devb, dev, 6451.8, max(3, min(args.steps, 30)), stem_scans=raw)
        print(json.dumps({k: round(v["us_per_scan"], 2) for k, v in fam.items()}), flush=True)
        return
    compute = torch.cuda.Stream(dev)
    copy = torch.cuda.Stream(dev)
    launches_per_step = 0
    with torch.cuda.stream(compute), torch.no_grad():
        for i in range(3):  # the plan cache learns its batches on the first scan: count a steady-state scan
            ops.reset_launch_count()
            hot.step(devb[i])
            launches_per_step = ops.launch_count()
        torch.cuda.synchronize()
        hot.scan_index = 0
        pipe = pipeline.ScanPipeline(hot, devb, use_graphs=use_graph, scans_in_flight=args.in_flight)
        outs = pipe.out

        # ---- device-resident throughput ("value") -------------------------------------------------------
        for i in range(args.warmup):
            pipe.submit()
        torch.cuda.synchronize()
        barrier(world)
        sampler = ClockSampler(local)
        if rank == 0:
            sampler.start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record(compute)
        for st in pipe.streams():
            st.wait_event(e0)
        for i in range(args.steps):
            pipe.submit()
        pipe.join(compute)
        e1.record(compute)
        torch.cuda.synchronize()
        barrier(world)
        ms_total = max_over_ranks(e0.elapsed_time(e1), world, dev)
        ms_step = ms_total / args.steps
        value = world * 1000.0 / ms_step

        # ---- end to end: host buffers in, labels out, copies inside the timed region --------------------
        def measure_e2e(pipe_, host_, devb_, nsteps, window=0):
            """`window` > 0: scan i also reads the raw-scan buffers of scans i-1 .. i-window (resident T-frame window):
            its graphs wait for their copies too, and a buffer is only overwritten once its later readers are done."""
            outs_ = pipe_.out
            h_labels = [torch.empty(args.points, dtype=torch.int64).pin_memory() for _ in range(N_SCANS)]
            h_sums = [torch.empty(stream.N_BOXES, 2, dtype=torch.int64).pin_memory() for _ in range(N_SCANS)]
            ready = [torch.cuda.Event() for _ in range(N_SCANS)]
            d2h = [torch.cuda.Event() for _ in range(N_SCANS)]
            sC = pipe_.streams()[-1]

            def e2e_loop(n):
                for i in range(n):
                    j = pipe_.submitted % N_SCANS
                    with torch.cuda.stream(copy):
                        # buffer j is free once the previous scan that used it has been voted on and read back
                        copy.wait_event(pipe_.m_done[j])
                        copy.wait_event(pipe_.v_done[j])
                        copy.wait_event(d2h[j])
                        for k in range(1, window + 1):  # the scans that read buffer j as an older frame
                            copy.wait_event(pipe_.m_done[(j + k) % N_SCANS])
                        if not os.environ.get("SMOS_E2E_NO_H2D"):  # experiment knob: how much of e2e is the copy
                            devb_[j].copy_from(host_[j])    # H2D of this scan's inputs (pinned -> HBM)
                        ready[j].record(copy)
                    pipe_.submit(ready[j], also_ready=[ready[(j - k) % N_SCANS] for k in range(1, window + 1)])
                    with torch.cuda.stream(sC):             # D2H of the step's result, behind the voting graph
                        if not os.environ.get("SMOS_E2E_NO_D2H"):
                            h_labels[j].copy_(outs_[j][0], non_blocking=True)
                            h_sums[j].copy_(outs_[j][1], non_blocking=True)
                        d2h[j].record(sC)

            for j in range(N_SCANS):
                d2h[j].record(sC)
            e2e_loop(max(4, min(args.warmup, 2 * N_SCANS)))
            torch.cuda.synchronize()
            barrier(world)
            f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            f0.record(compute)
            copy.wait_event(f0)
            for st in pipe_.streams():
                st.wait_event(f0)
            e2e_loop(nsteps)
            pipe_.join(compute)
            for j in range(N_SCANS):
                compute.wait_event(d2h[j])
            f1.record(compute)
            torch.cuda.synchronize()
            barrier(world)
            ms = max_over_ranks(f0.elapsed_time(f1), world, dev) / nsteps
            if os.environ.get("SMOS_E2E_PROFILE") and rank == 0:  # concurrency timeline of the same loop (CUPTI)
                import importlib.util
                from torch.profiler import ProfilerActivity, profile
                spec = importlib.util.spec_from_file_location(
                    "profile_pipeline", os.path.join(os.path.dirname(os.path.abspath(__file__)), "tools", "profile_pipeline.py"))
                pp = importlib.util.module_from_spec(spec)
                spec.loader.exec_module(pp)
                torch.cuda.synchronize()
                with profile(activities=[ProfilerActivity.CUDA]) as prof:
                    e2e_loop(32)
                    torch.cuda.synchronize()
                with open(os.environ["SMOS_E2E_PROFILE"], "a") as fh:
                    pp.analyse(prof, 32, out=fh)
            # host cost of submitting one scan (API calls of e2e_loop): a short burst that fits the launch queue, so the
            # host never waits for the device
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            e2e_loop(2 * N_SCANS)
            host_us = (time.perf_counter() - t0) / (2 * N_SCANS) * 1e6
            torch.cuda.synchronize()
            return {"steps": nsteps, "value": world * 1000.0 / ms, "unit": "scans/s", "ms_per_step": ms,
                    "host_submit_us_per_scan": max_over_ranks(host_us, world, dev),
                    "h2d_bytes_per_step": host_[0].nbytes(),
                    "d2h_bytes_per_step": h_labels[0].numel() * 8 + h_sums[0].numel() * 8}

        if args.e2e_only:
            hot_r = stream.HotPath(dev, n_points=args.points, seed=rank, point_major=not args.channel_major,
                                   vote_api=args.vote_api, branches=args.branches, batch_plans=args.explicit_plans)
            if os.environ.get("SMOS_E2E_ALL_FRAMES"):  # the previous definition: T aligned frames uploaded per scan
                host_r = [stream.make_host_raw_scan(rank * 1000 + i, args.points) for i in range(N_SCANS)]
                devb_r, win = [h.pack(device=dev) for h in host_r], 0
            else:
                host_r, _ = stream.make_host_resident_stream(rank, N_SCANS, args.points)
                devb_r, win = stream.link_window([h.pack(device=dev) for h in host_r]), 2
            torch.cuda.synchronize()
            pipe_r = pipeline.ScanPipeline(hot_r, devb_r, use_graphs=use_graph, scans_in_flight=args.in_flight)
            res = measure_e2e(pipe_r, host_r, devb_r, args.steps, window=win)
            if rank == 0:
                print(json.dumps(res), flush=True)
            return
        # (1) hot-path inputs themselves in host memory: the 64-channel point features (92 MB per scan) cross PCIe —
        #     something the reference never does (its PointNet stem produces them on the GPU)
        side_steps = min(args.steps, 300)
        e2e_feat = measure_e2e(pipe, host, devb, side_steps)
        e2e_feat["note"] = ("hot-path input tensors in pinned host memory (3x64xN point features = 92 MB of the H2D bytes): "
                            "PCIe bound; kept for reference")

        def hot_like():
            return stream.HotPath(dev, n_points=args.points, seed=rank, point_major=not args.channel_major,
                                  vote_api=args.vote_api, grids_channels_last=args.grids_channels_last,
                                  branches=args.branches, batch_plans=args.explicit_plans, ordered_gathers=not args.no_ordered_gathers,
                                  ordered_rv=args.ordered_rv, gather_taps=args.gather_taps)

        # (2) the LOADER's tensors in host memory, as in the reference (models/StreamMOS.py:86-103): 7-channel point
        #     features + BEV / range-view coordinates per frame; the PointNet stem (our fused kernel) runs on the device
        hot_l = hot_like()
        host_l = [stream.make_host_loader_scan(rank * 1000 + i, args.points) for i in range(N_SCANS)]
        devb_l = [h.pack(device=dev) for h in host_l]   # one flat buffer per scan: one H2D copy per step
        torch.cuda.synchronize()
        pipe_l = pipeline.ScanPipeline(hot_l, devb_l, use_graphs=use_graph, scans_in_flight=args.in_flight)
        e2e_loader = measure_e2e(pipe_l, host_l, devb_l, side_steps)
        e2e_loader["note"] = ("host buffers = the loader's output tensors (T x 7-channel point features, BEV and range-view "
                              "coordinates of every frame); stem + hot path on the device")
        del pipe_l, devb_l, hot_l, host_l
        # (3) headline e2e: the RAW scan in host memory — the loader's points before form_batch (range filtered, padded) and
        #     the range-view coordinates of the current frame; Quantize + make_point_feat (smos_form_batch, bit-exact), the
        #     PointNet stem, the whole hot path and the D2H of the labels are inside the timed region
        hot_r = hot_like()
        host_r = [stream.make_host_raw_scan(rank * 1000 + i, args.points) for i in range(N_SCANS)]
        devb_r = [h.pack(device=dev) for h in host_r]
        torch.cuda.synchronize()
        pipe_r = pipeline.ScanPipeline(hot_r, devb_r, use_graphs=use_graph, scans_in_flight=args.in_flight)
        e2e_frames = measure_e2e(pipe_r, host_r, devb_r, side_steps)
        e2e_frames["note"] = ("round-1 definition, kept for continuity: host buffers = T frames x N x (x, y, z, intensity) that "
                              "the HOST has pose-aligned, range filtered and padded + range-view coordinates of the current "
                              "frame + stand-ins; form_batch, stem, hot path and D2H on the device")
        del pipe_r, devb_r, hot_r, host_r
        # (4) headline e2e: the stream keeps the RAW scans of its T-frame window resident in HBM. Per scan the host hands
        #     over the new raw scan as read from the file (own sensor frame, unfiltered, unpadded), the window's pose_diff
        #     matrices, the range-view coordinates of the current frame and the stand-ins; pose alignment + range filter +
        #     padding of all T frames (smos_ingest_frames, bit-exact with the loader), Quantize + make_point_feat, the
        #     PointNet stem, the whole hot path and the D2H of the labels are inside the timed region
        hot_w = hot_like()
        host_w, _ = stream.make_host_resident_stream(rank, N_SCANS, args.points)
        devb_w = stream.link_window([h.pack(device=dev) for h in host_w])
        torch.cuda.synchronize()
        pipe_w = pipeline.ScanPipeline(hot_w, devb_w, use_graphs=use_graph, scans_in_flight=args.in_flight)
        e2e = measure_e2e(pipe_w, host_w, devb_w, args.steps, window=2)
        clocks = sampler.stop() if rank == 0 else None  # sampled across the timed regions
        e2e["note"] = ("host buffers = the NEW raw scan as read from the file (own sensor frame, no filter, no padding) + the "
                       "pose_diff matrices of the T = 3 window + range-view coordinates of the current frame (the loader's "
                       "SphereQuantize output, so that every input is bit-exact; sphere_quantize_on_device computes them on "
                       "the GPU too) + predicted labels and attention samples as stand-ins for network "
                       "intermediates; ONE H2D copy per scan. The two older raw scans stay resident in HBM; pose alignment, "
                       "range filter and padding of all T frames (smos_ingest_frames, bit-exact with the loader), Quantize + "
                       "make_point_feat, the PointNet stem, the whole hot path and the D2H of the labels are inside the "
                       "timed region; the copy of scan i+1 overlaps scan i")
        # (5) as (4), but nothing the loader computed crosses PCIe: SphereQuantize of the current frame on the device too
        #     (smos_sphere_quantize: floating point, within 1 ulp of the angle of numpy's float32 result — a handful of the
        #     120 k points change their range-view cell, which is why (4), bit-exact with the loader, stays the headline)
        hot_s = stream.HotPath(dev, n_points=args.points, seed=rank, point_major=not args.channel_major,
                               vote_api=args.vote_api, grids_channels_last=args.grids_channels_last, branches=args.branches,
                               batch_plans=args.explicit_plans, ordered_gathers=not args.no_ordered_gathers,
                               ordered_rv=args.ordered_rv, gather_taps=args.gather_taps, sphere_on_device=True)
        host_s, _ = stream.make_host_resident_stream(rank, N_SCANS, args.points, device_sphere=True)
        devb_s = stream.link_window([h.pack(device=dev) for h in host_s])
        torch.cuda.synchronize()
        pipe_s = pipeline.ScanPipeline(hot_s, devb_s, use_graphs=use_graph, scans_in_flight=args.in_flight)
        e2e_sphere = measure_e2e(pipe_s, host_s, devb_s, side_steps, window=2)
        e2e_sphere["note"] = ("as the headline, with SphereQuantize of the current frame on the device as well: the host hands "
                              "over the raw scan, the poses and the stand-ins only (floating-point range-view coordinates, "
                              "within 1 ulp of the angle of the loader's)")
        del pipe_s, devb_s, hot_s, host_s
        e2e["sphere_quantize_on_device"] = e2e_sphere
        e2e["raw_scan_all_frames_from_host"] = e2e_frames
        e2e["loader_tensors"] = e2e_loader
        e2e["hot_path_inputs_over_pcie"] = e2e_feat
        del pipe_w, devb_w, hot_w

        # ---- dominant kernel: the dense writer of VoxelMaxPool #1 (3 x 64 x 512 x 512 fp32 out) --------------
        # every stage of the call runs once, then the WRITE stage alone is re-launched and timed with CUDA
        # events on its stream; the 201 MB output per launch (> 126 MB L2) is its own cache flush
        ab = stream.algorithmic_bytes(args.points)
        ksteps = min(args.steps, 200)
        plan1 = ops.pool_plan(devb[0].coord_bev, (512, 512), (1.0, 1.0))
        ws1 = ops.pool_workspace(3, 64, args.points, dev)
        out1 = torch.empty(3, 64, 512, 512, device=dev)
        ops.voxel_maxpool_forward(devb[0].feat, plan1, out=out1, workspace=ws1)
        for i in range(3):
            ops.voxel_maxpool_forward(devb[0].feat, plan1, out=out1, stages=ops.POOL_STAGE_WRITE, workspace=ws1)
        k0 = [torch.cuda.Event(enable_timing=True) for _ in range(ksteps)]
        k1 = [torch.cuda.Event(enable_timing=True) for _ in range(ksteps)]
        for i in range(ksteps):
            k0[i].record(compute)
            ops.voxel_maxpool_forward(devb[0].feat, plan1, out=out1, stages=ops.POOL_STAGE_WRITE, workspace=ws1)
            k1[i].record(compute)
        torch.cuda.synchronize()
        kern_ms = sum(a.elapsed_time(b) for a, b in zip(k0, k1)) / ksteps
        # whole VoxelMaxPool #1 (plan + permute + reduce + combine + write) for the op-level figure
        for i in range(3):
            deep_point_pool1(devb[i])
        p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        p0.record(compute)
        for i in range(ksteps):
            deep_point_pool1(devb[i % N_SCANS])
        p1.record(compute)
        torch.cuda.synchronize()
        pool1_ms = p0.elapsed_time(p1) / ksteps
        breakdown = op_breakdown(hot, devb, compute, min(args.steps, 50)) if (rank == 0) else None
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        families = family_breakdown(hot, devb, dev, peak, max(3, min(args.steps, 30))) if rank == 0 else None

        # ---- variant (reported beside the headline, not instead of it): what a channels_last model and the fused
        # streaming voting API buy on the same workload ------------------------------------------
        variants = None
        if rank == 0 and world == 1 and not args.no_variants and use_graph:
            del pipe

            def measure_variant(**kw):
                base = dict(n_points=args.points, seed=rank, point_major=not args.channel_major, vote_api=args.vote_api,
                            grids_channels_last=args.grids_channels_last, branches=args.branches,
                            batch_plans=args.explicit_plans, ordered_gathers=not args.no_ordered_gathers,
                            ordered_rv=args.ordered_rv, gather_taps=args.gather_taps)
                base.update(kw)
                hot2 = stream.HotPath(dev, **base)
                pipe2 = pipeline.ScanPipeline(hot2, devb, use_graphs=True, scans_in_flight=args.in_flight)
                vsteps = min(args.steps, 400)
                for i in range(20):
                    pipe2.submit()
                torch.cuda.synchronize()
                v0, v1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                v0.record(compute)
                for st in pipe2.streams():
                    st.wait_event(v0)
                for i in range(vsteps):
                    pipe2.submit()
                pipe2.join(compute)
                v1.record(compute)
                torch.cuda.synchronize()
                vms = v0.elapsed_time(v1) / vsteps
                del pipe2, hot2
                return {"value": 1000.0 / vms, "unit": "scans/s", "ms_per_step": vms}

            variants = {}
            if not (args.grids_channels_last and args.vote_api == "fused"):
                variants["channels_last_cnn_grids+fused_voting_api"] = dict(
                    measure_variant(vote_api="fused", grids_channels_last=True),
                    note="same kernels; gathers read channels_last feature maps (model.to(memory_format=channels_last)) and "
                         "voting takes float xyz + uint8 labels (smos_vote_fused) instead of the int64 staging of "
                         "voxel_voting.py:234-243; results identical (tests)")
            variants["explicit_plan_api" if not args.explicit_plans else "reference_signatures"] = dict(
                measure_variant(batch_plans=not args.explicit_plans),
                note="same workload through the other operator API (explicit: five plans per batch of launches, plan= / order=; "
                     "reference: the reference's arguments only + plan cache)")
            if not args.branches:
                variants["parallel_branches"] = dict(
                    measure_variant(branches=True, batch_plans=True),
                    note="experiment: pool #1 | half-scale chain | quarter-scale chain | gather #5 as parallel graph branches — "
                         "possible only because the CNN outputs are resident stand-ins; NOT the reference's data flow")

    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6.65 TB/s (B200_PROFILING.md)"
    # algorithmic bytes of the writer per launch: the dense output (4*B'*C*H*W) + count/start per cell (8*B'*H*W);
    # the rows of occupied cells are < 10 % more and are not counted
    k_bytes = 4 * 3 * 64 * 512 * 512 + 8 * 3 * 512 * 512
    achieved = k_bytes / (kern_ms * 1e-3) / 1e9
    traffic = None
    try:  # dram__bytes_read.sum + dram__bytes_write.sum of this kernel from the committed ncu capture
        traffic = json.load(open(os.path.join(ROOT, "profiles", "dominant_kernel_traffic.json")))["bytes_per_launch"]
    except Exception:
        pass
    roofline = {"bound": "hbm", "kernel": "pool_write_kernel (dense writer of VoxelMaxPool #1: 3x64x512x512 fp32; the largest "
                                          "single byte mover of the path, 30 % of its algorithmic bytes)",
                "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": k_bytes, "kernel_ms": kern_ms, "traffic": traffic,
                "voxelmaxpool1_op": {"algorithmic_bytes": ab["pool"][0], "ms": pool1_ms,
                                     "achieved": ab["pool"][0] / (pool1_ms * 1e-3) / 1e9,
                                     "frac": ab["pool"][0] / (pool1_ms * 1e-3) / 1e9 / peak,
                                     "note": "eager launches: plan (4 kernels) + permute + reduce + combine + write"},
                "whole_path": {"algorithmic_bytes_per_scan": ab["total"],
                               "achieved": ab["total"] / (ms_step * 1e-3) / 1e9,
                               "frac": ab["total"] / (ms_step * 1e-3) / 1e9 / peak,
                               "note": "pipelined wall time per scan (ms_per_step): kernels of %d scans in flight overlap" % args.in_flight}}
    if families:
        # SURVEY 8d's own formula: algorithmic bytes / SUM of the hot-path kernel time (no overlap between families)
        sum_us = sum(f["us_per_scan"] for f in families.values())
        roofline["per_family"] = families
        roofline["whole_path_sum_of_kernels"] = {"us_per_scan": sum_us, "algorithmic_bytes_per_scan": ab["total"],
                                                 "achieved": ab["total"] / (sum_us * 1e-6) / 1e9,
                                                 "frac": ab["total"] / (sum_us * 1e-6) / 1e9 / peak}
        with_bytes = {k: f for k, f in families.items() if f.get("frac") is not None and k != "msda"}
        lim = min(with_bytes, key=lambda k: with_bytes[k]["frac"])
        roofline["limiter"] = {"family": lim, "frac": with_bytes[lim]["frac"],
                               "note": "lowest fraction among the HBM-sized families (msda is 5 MB, L2 resident, latency bound); "
                                       "`kernel` above is the single largest byte mover, not the limiter"}
    if rank != 0:
        return
    line = {"metric": METRIC, "value": value, "unit": "scans/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": dict(workload_config(args, use_graph, world),
                           host_affinity=("rank pinned to %d cores of its GPU's NUMA node" % len(pinned_cpus))
                           if pinned_cpus else "unchanged"),
            "roofline": roofline, "e2e": e2e,
            "gpu_launches": launches_per_step * args.steps, "gpu_launches_per_step": launches_per_step,
            "clocks": clocks, "breakdown_ms": breakdown, "variants": variants,
            "reference_signature": None if args.explicit_plans else {
                "value": value, "unit": "scans/s",
                "note": "`value` IS the reference-signature leg: every operator is called with the reference's arguments "
                        "only; variants.explicit_plan_api is the same workload through plan= / order="}}
    if cpu_hot_state is not None:
        scans = [h for h in host[:4]]
        sps, ms, cores, done = time_cpu_path(cpu_hot_state, scans, args.cpu_scans, budget_s=30.0)
        line["cpu_baseline"] = {"value": sps, "unit": "scans/s", "cores": cores, "kind": "port", "ms_per_scan": ms,
                                "sample": "%d scans of the same workload, oracle/cpu_path.py (torch CPU ops)" % done}
    else:
        line["cpu_baseline"] = None
    print(json.dumps(line), flush=True)
    if args.breakdown and breakdown:
        for k, v in breakdown.items():
            sys.stderr.write("%-28s %8.4f ms\n" % (k, v))


def family_breakdown(hot, devb, dev, peak, iters, stem_scans=None):
    """Per-family device time of the hot path (roofline = algorithmic bytes / sum of hot-path kernel time).

    Each family's launches for all resident scans are captured into ONE CUDA graph (no host launch gaps) and the
    graph is replayed `iters` times between two CUDA events on its stream; the ~100 MB of inputs per scan x 8 scans
    rotate through L2 as in the stream. Consumers get pre-built plans, so plan building is its own family."""
    import torch
    from streammos_b200 import MultiScaleDeformableAttention as MSDA
    from streammos_b200 import deep_point, ops, stream
    ab = stream.algorithmic_bytes(hot.n_points)
    nb = len(devb)
    specs = lambda b: [(b.coord_bev, (512, 512), (1.0, 1.0)), (b.coord_rv, (32, 1024), (0.5, 0.5)),
                       (b.coord_bev[:1], (256, 256), (0.5, 0.5)), (b.coord_rv, (16, 512), (0.25, 0.25)),
                       (b.coord_bev[:1], (128, 128), (0.25, 0.25))]
    st = torch.cuda.Stream(dev)
    res = {}
    with torch.cuda.stream(st), torch.no_grad():
        P = [[ops.pool_plan(*sp) for sp in specs(b)] for b in devb]
        keep = []
        for b, pl in zip(devb, P):  # intermediates of the cascade, once per scan
            cur, rv = b.coord_bev[:1], b.coord_rv
            x0_pt = hot.g_half(hot.x0, cur, pl[2])
            x0_rv = deep_point.VoxelMaxPool(x0_pt, rv, (32, 1024), (0.5, 0.5), pl[1])
            x0_pt2 = hot.g_half(x0_rv, rv, pl[1])
            x1_pt = hot.g_quarter(hot.x1, cur, pl[4])
            x1_rv = deep_point.VoxelMaxPool(x1_pt, rv, (16, 512), (0.25, 0.25), pl[3])
            x1_pt2 = hot.g_quarter(x1_rv, rv, pl[3])
            keep.append((x0_pt, x0_rv, x0_pt2, x1_pt, x1_rv, x1_pt2))
        value = hot.memory.view(1, stream.MEM_HW * stream.MEM_HW, stream.N_HEADS, stream.HEAD_DIM)

        def f_plans(j):
            if hot.batch_plans:
                return ops.pool_plan_multi(specs(devb[j]))
            # reference signatures: the plan cache builds what the cascade asks for, in the cascade's order — one by one
            # on the first scan, then in the batches it learned (plan_cache.py); inside this capture every scan's
            # coordinate tensors are new to the cache, as in the stream
            sp = specs(devb[j])
            return [ops.cached_pool_plan(*sp[i]) for i in (0, 2, 1, 4, 3)]

        raw_feat = None
        if stem_scans is not None:  # raw-scan path: the stem's output (point-major rows) is what pool #1 reads
            from streammos_b200 import synthetic
            stem_args = (synthetic.RANGE_X, synthetic.RANGE_Y, synthetic.RANGE_Z, hot.size) + tuple(hot.stem.fused_parameters())
            raw_feat = [ops.point_stem_forward_raw(r.points, *stem_args, point_major_out=hot.point_major)[0] for r in stem_scans]

        def f_stem(j):
            return ops.point_stem_forward_raw(stem_scans[j].points, *stem_args, point_major_out=hot.point_major)

        def f_pool1(j):
            feat = raw_feat[j] if raw_feat is not None else devb[j].feat
            return deep_point.VoxelMaxPool(feat, devb[j].coord_bev, (512, 512), (1.0, 1.0), P[j][0])

        def f_pools(j):
            b, k, pl = devb[j], keep[j], P[j]
            cur, rv = b.coord_bev[:1], b.coord_rv
            return (deep_point.VoxelMaxPool(k[0], rv, (32, 1024), (0.5, 0.5), pl[1]),
                    deep_point.VoxelMaxPool(k[2], cur, (256, 256), (0.5, 0.5), pl[2]),
                    deep_point.VoxelMaxPool(k[3], rv, (16, 512), (0.25, 0.25), pl[3]),
                    deep_point.VoxelMaxPool(k[5], cur, (128, 128), (0.25, 0.25), pl[4]))

        def f_gathers(j):
            b, k, pl = devb[j], keep[j], P[j]
            cur, rv = b.coord_bev[:1], b.coord_rv
            return (hot.g_half(hot.x0, cur, pl[2]), hot.g_half(k[1], rv, pl[1]), hot.g_quarter(hot.x1, cur, pl[4]),
                    hot.g_quarter(k[4], rv, pl[3]), hot.g_half(hot.dec, cur, pl[2]))

        def f_msda(j):
            b = devb[j]
            h = MSDA.ms_deform_attn_forward(value, hot.shapes, hot.lsi, b.loc[0], b.attn[0], 256)
            return MSDA.ms_deform_attn_forward(h.view_as(value), hot.shapes, hot.lsi, b.loc[1], b.attn[1], 256)

        def f_vote(j):
            hot.scan_index = j
            return hot.long_term_voting(devb[j])

        fams = ([("stem (raw-scan path only)", f_stem, 0)] if stem_scans is not None else []) + \
               [("plans", f_plans, 0), ("pool1", f_pool1, ab["pool"][0]), ("pools2-5", f_pools, sum(ab["pool"][1:])),
                ("gathers", f_gathers, sum(ab["gather"])), ("msda", f_msda, ab["msda"]), ("voting", f_vote, ab["vote"])]
        for name, fn, nbytes in fams:
            for j in range(nb):
                fn(j)
            torch.cuda.synchronize(dev)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=st):
                outs = [fn(j) for j in range(nb)]
            for _ in range(2):
                g.replay()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(st)
            for _ in range(iters):
                g.replay()
            e1.record(st)
            torch.cuda.synchronize(dev)
            ms = e0.elapsed_time(e1) / (iters * nb)
            ent = {"us_per_scan": ms * 1e3, "algorithmic_bytes": nbytes}
            if nbytes:
                ent["achieved"] = nbytes / (ms * 1e-3) / 1e9
                ent["frac"] = ent["achieved"] / peak
            res[name] = ent
            del g, outs
        hot.scan_index = 0
    return res


def deep_point_pool1(b):
    from streammos_b200 import deep_point
    return deep_point.VoxelMaxPool(b.feat, b.coord_bev, (512, 512), (1.0, 1.0))


def op_breakdown(hot, devb, stream_, iters):
    """CUDA-event time of each operator of the step (eager launches, same inputs) — explains `value`."""
    import torch
    from streammos_b200 import MultiScaleDeformableAttention as MSDA
    from streammos_b200 import deep_point, ops, stream, voting
    res = {}

    def timeit(name, fn):
        for _ in range(3):
            fn(0)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream_)
        for i in range(iters):
            fn(i)
        b.record(stream_)
        torch.cuda.synchronize()
        res[name] = a.elapsed_time(b) / iters

    B = devb
    nb = len(B)
    cur = lambda i: B[i % nb].coord_bev[:1]
    rv = lambda i: B[i % nb].coord_rv
    timeit("pool1_bev512_c64x3", lambda i: deep_point.VoxelMaxPool(B[i % nb].feat, B[i % nb].coord_bev, (512, 512), (1.0, 1.0)))
    x0p = hot.g_half(hot.x0, cur(0))
    x1p = hot.g_quarter(hot.x1, cur(0))
    timeit("gather1_x0_256_c32", lambda i: hot.g_half(hot.x0, cur(i)))
    timeit("pool2_rv32x1024_c32", lambda i: deep_point.VoxelMaxPool(x0p, rv(0), (32, 1024), (0.5, 0.5)))
    x0rv = deep_point.VoxelMaxPool(x0p, rv(0), (32, 1024), (0.5, 0.5))
    timeit("gather2_rv32x1024_c32", lambda i: hot.g_half(x0rv, rv(i)))
    timeit("pool3_bev256_c32", lambda i: deep_point.VoxelMaxPool(x0p, cur(0), (256, 256), (0.5, 0.5)))
    timeit("gather3_x1_128_c64", lambda i: hot.g_quarter(hot.x1, cur(i)))
    timeit("pool4_rv16x512_c64", lambda i: deep_point.VoxelMaxPool(x1p, rv(0), (16, 512), (0.25, 0.25)))
    x1rv = deep_point.VoxelMaxPool(x1p, rv(0), (16, 512), (0.25, 0.25))
    timeit("gather4_rv16x512_c64", lambda i: hot.g_quarter(x1rv, rv(i)))
    timeit("pool5_bev128_c64", lambda i: deep_point.VoxelMaxPool(x1p, cur(0), (128, 128), (0.25, 0.25)))
    timeit("gather5_dec_256_c64", lambda i: hot.g_half(hot.dec, cur(i)))
    value = hot.memory.view(1, 4096, 4, 32)
    timeit("msda_fwd_x2", lambda i: (MSDA.ms_deform_attn_forward(value, hot.shapes, hot.lsi, B[i % nb].loc[0], B[i % nb].attn[0], 256),
                                     MSDA.ms_deform_attn_forward(value, hot.shapes, hot.lsi, B[i % nb].loc[1], B[i % nb].attn[1], 256)))
    timeit("voting_voxel+instance", lambda i: hot.long_term_voting(B[i % nb]))
    plan = ops.pool_plan(B[0].coord_bev, (512, 512), (1.0, 1.0))
    timeit("pool1_plan_only", lambda i: ops.pool_plan(B[i % nb].coord_bev, (512, 512), (1.0, 1.0)))
    out1 = torch.empty(3, 64, 512, 512, device=hot.device)
    timeit("pool1_forward_only", lambda i: ops.voxel_maxpool_forward(B[i % nb].feat, plan, out=out1))
    return res


def main():
    args = parse_args()
    if args.impl == "reference":  # CPU only: rank 0 runs it, the other ranks exit without work (no process group)
        run_reference(args, int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")))
        return
    world, rank, local = dist_setup()
    try:
        if True:
            run_b200(args, world, rank, local)
    finally:
        if world > 1:
            import torch.distributed as dist
            if dist.is_initialized():
                dist.destroy_process_group()


if __name__ == "__main__":
    main()
