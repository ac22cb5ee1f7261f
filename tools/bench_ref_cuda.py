"""Same-box bar: the REFERENCE's own CUDA kernels (compiled unmodified for sm_100a by tools/install_ref.py; deformattn with the
2-token torch-2 patch) timed per operator next to this repo's kernels, same inputs, same B200, CUDA events.

    python tools/bench_ref_cuda.py [--iters 30] [--points 120000]

Reference side = what the reference model executes per call: deep_point/__init__.py (2 fills + 3 metadata uploads + 3
kernels) through its compiled point_deep.cuda_kernel; networks/backbone.py BilinearSample (4 elementwise kernels + stack +
F.grid_sample); deformattn MSDeformAttnFunction through its compiled extension. Ours = the drop-in modules with the
reference's arguments (plan cache warm as inside a scan: pools and gathers of one scan share plans; the plan-build cost
is listed on its own line). Also the whole model: AttNet.infer scans/s with the reference's CUDA ops vs with the drop-in.
"""
import argparse
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import refmodel  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--iters", type=int, default=30)
ap.add_argument("--points", type=int, default=120000)
ap.add_argument("--model-scans", type=int, default=12)
ap.add_argument("--model-points", type=int, default=120000)
a = ap.parse_args()
dev = torch.device("cuda:0")
N = a.points


def timeit(fn, iters=a.iters):
    for i in range(3):
        fn(i)
    torch.cuda.synchronize()
    e0 = [torch.cuda.Event(enable_timing=True) for _ in range(iters)]
    e1 = [torch.cuda.Event(enable_timing=True) for _ in range(iters)]
    for i in range(iters):
        e0[i].record()
        fn(i)
        e1[i].record()
    torch.cuda.synchronize()
    ts = sorted(x.elapsed_time(y) * 1e3 for x, y in zip(e0, e1))
    return ts[len(ts) // 2]


def graph_time(fn, iters=a.iters):
    """Device time of fn() without host launch gaps: captured once, replayed."""
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        for _ in range(3):
            fn(0)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        try:
            with torch.cuda.graph(g, stream=s):
                fn(0)
        except Exception:
            torch.cuda.synchronize()
            return None
        for _ in range(3):
            g.replay()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(s)
        for _ in range(iters):
            g.replay()
        e1.record(s)
        torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / iters


torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
from streammos_b200 import plan_cache, stream  # noqa: E402
scans = [stream.make_host_scan(i, N).to(dev) for i in range(4)]
g = torch.Generator().manual_seed(0)
x0 = torch.randn(1, 32, 256, 256, generator=g).relu_().to(dev)
x1 = torch.randn(1, 64, 128, 128, generator=g).relu_().to(dev)
dec = torch.randn(1, 64, 256, 256, generator=g).relu_().to(dev)
x0rv = torch.randn(1, 32, 32, 1024, generator=g).relu_().to(dev)
x1rv = torch.randn(1, 64, 16, 512, generator=g).relu_().to(dev)
f32 = torch.randn(1, 32, N, 1, generator=g).relu_().to(dev)
f64 = torch.randn(1, 64, N, 1, generator=g).relu_().to(dev)
value = torch.randn(1, 4096, 4, 32, generator=g).to(dev)
shapes = torch.tensor([[64, 64]], dtype=torch.int64, device=dev)
lsi = torch.zeros(1, dtype=torch.int64, device=dev)
gout = torch.randn(1, 4096, 128, generator=g).to(dev)

rows = []


def collect(mode):
    """-> dict op -> (eager us, graph us or None) for one operator set."""
    refmodel.purge()
    refmodel._stub_third_party()
    sys.path.insert(0, refmodel.ref_root())
    import types
    if mode == "reference":
        pkg = types.ModuleType("point_deep")
        pkg.cuda_kernel = refmodel._load_so("ref_point_deep_cuda", refmodel.ref_ext("point_deep_cuda"))
        pkg.cpu_kernel = types.ModuleType("point_deep.cpu_kernel")
        sys.modules["point_deep"] = pkg
        sys.modules["point_deep.cuda_kernel"] = pkg.cuda_kernel
        sys.modules["point_deep.cpu_kernel"] = pkg.cpu_kernel
        sys.modules["MultiScaleDeformableAttention"] = refmodel._load_so("ref_msda", refmodel.ref_ext("msda"))
    else:
        from streammos_b200 import dropin
        dropin.install()
    import deep_point
    import networks.backbone as bb
    if mode != "reference":
        from streammos_b200 import dropin
        dropin.install()
    from deformattn.functions import MSDeformAttnFunction
    gh, gq = bb.BilinearSample(32, (0.5, 0.5)), bb.BilinearSample(64, (0.25, 0.25))
    S = lambda i: scans[i % 4]
    cur = lambda i: S(i).coord_bev[:1].contiguous()
    curs = [cur(i) for i in range(4)]
    rvs = [S(i).coord_rv.contiguous() for i in range(4)]
    bevs = [S(i).coord_bev.contiguous() for i in range(4)]
    ops = {
        "pool1 3x64xN -> 512^2": lambda i: deep_point.VoxelMaxPool(S(i).feat, bevs[i % 4], (512, 512), (1.0, 1.0)),
        "pool2 32ch -> rv 32x1024": lambda i: deep_point.VoxelMaxPool(f32, rvs[i % 4], (32, 1024), (0.5, 0.5)),
        "pool3 32ch -> bev 256^2": lambda i: deep_point.VoxelMaxPool(f32, curs[i % 4], (256, 256), (0.5, 0.5)),
        "pool4 64ch -> rv 16x512": lambda i: deep_point.VoxelMaxPool(f64, rvs[i % 4], (16, 512), (0.25, 0.25)),
        "pool5 64ch -> bev 128^2": lambda i: deep_point.VoxelMaxPool(f64, curs[i % 4], (128, 128), (0.25, 0.25)),
        "gather1 32ch@256^2": lambda i: gh(x0, curs[i % 4]),
        "gather2 32ch@32x1024": lambda i: gh(x0rv, rvs[i % 4]),
        "gather3 64ch@128^2": lambda i: gq(x1, curs[i % 4]),
        "gather4 64ch@16x512": lambda i: gq(x1rv, rvs[i % 4]),
        "gather5 64ch@256^2": lambda i: gh(dec, curs[i % 4]),
        "msda fwd (1 layer)": lambda i: MSDeformAttnFunction.apply(value, shapes, lsi, S(i).loc[0], S(i).attn[0], 256),
    }
    res = {}
    with torch.no_grad():
        for name, fn in ops.items():
            plan_cache.clear()
            if mode != "reference":  # warm the cache for every coordinate tensor: inside a scan the plans are shared
                for i in range(4):
                    fn(i)
            res[name] = (timeit(fn), graph_time(fn))
        if mode != "reference":
            from streammos_b200 import ops as b200ops
            res["plan builds of one scan (5 plans, 3 batches)"] = (None, graph_time(lambda i: (
                plan_cache.clear() or True) and [b200ops.pool_plan(bevs[0], (512, 512), (1.0, 1.0)),
                                                b200ops.pool_plan_multi([(curs[0], (256, 256), (0.5, 0.5)), (curs[0], (128, 128), (0.25, 0.25))]),
                                                b200ops.pool_plan_multi([(rvs[0], (32, 1024), (0.5, 0.5)), (rvs[0], (16, 512), (0.25, 0.25))])]))
    v = value.clone().requires_grad_(True)
    loc = scans[0].loc[0].clone().requires_grad_(True)
    at = scans[0].attn[0].clone().requires_grad_(True)

    def bwd(i):
        out = MSDeformAttnFunction.apply(v, shapes, lsi, loc, at, 256)
        out.backward(gout)
    res["msda fwd+bwd (1 layer, autograd)"] = (timeit(bwd), None)
    return res


ref = collect("reference")
ours = collect("b200")
print("%-46s %12s %12s %12s %12s %8s" % ("operator (N=%d)" % N, "ref eager us", "ref graph us", "b200 eager", "b200 graph", "x (dev)"))
for k in ours:
    r, o = ref.get(k, (None, None)), ours[k]
    f = lambda x: "%12.1f" % x if x is not None else "%12s" % "-"
    dev_r = r[1] if r[1] is not None else r[0]
    dev_o = o[1] if o[1] is not None else o[0]
    ratio = "%8.1f" % (dev_r / dev_o) if (dev_r and dev_o) else "%8s" % "-"
    print("%-46s %s %s %s %s %s" % (k, f(r[0]), f(r[1]), f(o[0]), f(o[1]), ratio))
print("(eager = CUDA events around one Python call, includes host launch gaps; graph = same call captured and replayed: "
      "device time only; the reference's deep_point/__init__.py uploads metadata with synchronous copies and cannot be captured)")

# ---- the whole model: AttNet.infer scans/s, reference CUDA ops vs drop-in ------------------------------------------
if a.model_scans > 0:
    batches = [refmodel.make_batch(500 + i, a.model_points) for i in range(4)]
    for mode in ("cuda_reference", "b200"):
        net, d = refmodel.load_attnet(mode, seed=0)
        bs = [{k: v.to(d) for k, v in b.items()} for b in batches]
        store = None
        with torch.no_grad():
            for i in range(3):
                store = net.infer(bs[i % 4], i, store)[-1]
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for i in range(a.model_scans):
                store = net.infer(bs[i % 4], 3 + i, store)[-1]
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
        print("AttNet.infer (unmodified reference model, B=1, T=3, N=%d, fp32, eager): %-15s %7.2f ms/scan  %6.1f scans/s"
              % (a.model_points, mode, dt / a.model_scans * 1e3, a.model_scans / dt))
        refmodel.purge()
