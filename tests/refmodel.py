"""Test infrastructure: load the UNMODIFIED reference model (`models/StreamMOS.py:AttNet`, which pulls in
`networks/multi_view_encoder.py`, `networks/backbone.py`, `deformattn/`) from baseline/_ref/StreamMOS (installed by
tools/install_ref.py; /root/reference itself is only read in the build container) on top of one of four operator sets:

  "cpu_reference"  the reference's own deep_point/__init__.py over its CPU kernel (oracle/_ref, compiled unmodified),
                   its BilinearSample (F.grid_sample) and ms_deform_attn_core_pytorch — CPU tensors
  "cuda_reference" the reference's own CUDA extensions compiled for sm_100a (baseline/_ref/ext; deformattn with the
                   2-token torch-2 patch), its BilinearSample — the same-box GPU baseline
  "torch_gpu"      torch-native restatements on CUDA tensors (scatter_reduce amax pooling, F.grid_sample,
                   ms_deform_attn_core_pytorch): isolates our operators from CPU-vs-cuDNN convolution differences
  "b200"           streammos_b200.dropin.install(): the product

Only tests/ and tools/ import this file.
"""
import importlib
import importlib.util
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_PACKAGES = ("models", "networks", "deformattn", "deep_point", "point_deep", "MultiScaleDeformableAttention", "utils",
                "config", "datasets")


def ref_root():
    for p in (os.path.join(ROOT, "baseline", "_ref", "StreamMOS"), "/root/reference"):
        if os.path.isdir(os.path.join(p, "models")):
            return p
    return None


def ref_ext(name):
    p = {"point_deep_cuda": os.path.join(ROOT, "baseline", "_ref", "ext", "point_deep_cuda", "ref_point_deep_cuda.so"),
         "msda": os.path.join(ROOT, "baseline", "_ref", "ext", "msda", "ref_msda.so")}[name]
    return p if os.path.exists(p) else None


def _load_so(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def purge():
    """Forget every module of the reference tree (and the names the drop-in registers)."""
    for k in list(sys.modules):
        if k.split(".")[0] in REF_PACKAGES:
            del sys.modules[k]
    root = ref_root()
    while root in sys.path:
        sys.path.remove(root)


def _stub_third_party():
    if "pytz" not in sys.modules:  # utils/logger.py imports it at module level; not installed here, not on the path
        try:
            import pytz  # noqa: F401
        except ImportError:
            pz = types.ModuleType("pytz")
            pz.utc = None
            pz.timezone = lambda *a, **k: None
            sys.modules["pytz"] = pz


def _core_pytorch_function(func_module):
    class _Fn:
        @staticmethod
        def apply(value, shapes, lsi, loc, attn, step):
            return func_module.ms_deform_attn_core_pytorch(value, shapes, loc, attn)
    return _Fn


def _patch_msda_function(fn):
    import deformattn.functions as DF
    import deformattn.functions.ms_deform_attn_func as F_
    import deformattn.modules.ms_deform_attn as MM
    F_.MSDeformAttnFunction = fn
    DF.MSDeformAttnFunction = fn
    MM.MSDeformAttnFunction = fn


def _torch_voxel_maxpool(pcds_feat, pcds_ind, output_size, scale_rate):
    """torch-native VoxelMaxPool on any device (the "scatter_reduce equivalent" of BASELINE.json), same index rule."""
    B, C, N, _ = pcds_feat.shape
    H, W = int(output_size[0]), int(output_size[1])
    ih = (pcds_ind[:, :, 0, 0] * float(scale_rate[0])).to(torch.int64)
    iw = (pcds_ind[:, :, 1, 0] * float(scale_rate[1])).to(torch.int64)
    ok = (ih >= 0) & (ih < H) & (iw >= 0) & (iw < W)
    cell = torch.where(ok, ih * W + iw, torch.full_like(ih, H * W))
    out = torch.zeros(B, C, H * W + 1, dtype=pcds_feat.dtype, device=pcds_feat.device)
    out.scatter_reduce_(2, cell[:, None, :].expand(B, C, N), pcds_feat[..., 0], "amax", include_self=False)
    return out[:, :, : H * W].reshape(B, C, H, W)


def load_attnet(mode, seed=0, state_dict=None):
    """-> (net, device). The reference model with random weights (seeded) or `state_dict`, in eval mode."""
    root = ref_root()
    if root is None:
        raise RuntimeError("reference tree not installed (python tools/install_ref.py)")
    purge()
    _stub_third_party()
    sys.path.insert(0, root)
    device = torch.device("cpu" if mode == "cpu_reference" else "cuda")
    if mode == "b200":
        from streammos_b200 import dropin
        dropin.install()
    else:
        sys.modules["MultiScaleDeformableAttention"] = types.ModuleType("MultiScaleDeformableAttention")
        pkg = types.ModuleType("point_deep")
        stub = types.ModuleType("point_deep.cuda_kernel")
        pkg.cpu_kernel, pkg.cuda_kernel = stub, stub
        if mode == "cpu_reference":
            sys.path.insert(0, ROOT)
            from oracle import build_ref
            cpu = build_ref.load()
            if cpu is None:
                raise RuntimeError("oracle/_ref was not built")
            pkg.cpu_kernel = cpu
        if mode == "cuda_reference":
            pkg.cuda_kernel = _load_so("ref_point_deep_cuda", ref_ext("point_deep_cuda"))
            sys.modules["MultiScaleDeformableAttention"] = _load_so("ref_msda", ref_ext("msda"))
        sys.modules["point_deep"] = pkg
        sys.modules["point_deep.cpu_kernel"] = pkg.cpu_kernel
        sys.modules["point_deep.cuda_kernel"] = pkg.cuda_kernel
        import deformattn.functions.ms_deform_attn_func as F_
        if mode in ("cpu_reference", "torch_gpu"):
            _patch_msda_function(_core_pytorch_function(F_))
        if mode == "torch_gpu":
            import deep_point
            deep_point.VoxelMaxPool = _torch_voxel_maxpool
    if mode == "b200":  # dropin.install() ran before `networks` was importable: patch the classes now
        import networks.backbone  # noqa: F401
        from streammos_b200 import dropin
        dropin.install()
    import config.StreamMOS as CFG
    from models import StreamMOS as SM
    _, _, model_param, _ = CFG.get_config()
    torch.manual_seed(seed)
    net = SM.AttNet(model_param)
    if state_dict is not None:
        net.load_state_dict(state_dict)
    else:
        randomize_norm_stats(net, seed)
    net.eval().to(device)
    return net, device


def randomize_norm_stats(net, seed):
    """Fresh BatchNorms (mean 0, var 1, weight 1) make every layer of a random network the same scale; non-trivial
    running statistics and affine parameters exercise the BatchNorm folding of the fused stem as a checkpoint would."""
    g = torch.Generator().manual_seed(seed + 1)
    with torch.no_grad():
        for m in net.modules():
            if isinstance(m, torch.nn.BatchNorm2d):
                m.weight.copy_(torch.rand(m.num_features, generator=g) * 0.5 + 0.75)
                m.bias.copy_(torch.randn(m.num_features, generator=g) * 0.1)
                m.running_mean.copy_(torch.randn(m.num_features, generator=g) * 0.1)
                m.running_var.copy_(torch.rand(m.num_features, generator=g) * 0.5 + 0.75)
        bn0 = net.point_pre.layer[0].layer[0]  # the input BatchNorm sees raw metres
        bn0.running_mean.copy_(torch.tensor([0.5, -0.3, -1.2, 0.3, 18.0, 0.5, 0.5]))
        bn0.running_var.copy_(torch.tensor([300.0, 280.0, 0.8, 0.05, 150.0, 0.08, 0.08]))


def make_batch(seed, n_points, views=1, t_frames=3):
    """One synthetic val-loader batch (datasets/data_StreamMOS.py:565-574 shapes, leading batch dim of 1 that
    AttNet.infer squeezes): pcds_xyzi (1, BS, T, 7, N, 1), pcds_coord (1, BS, T, N, 3, 1), pcds_sphere_coord
    (1, BS, T, N, 2, 1). views > 1 repeats the scan with the TTA flips of :495-513 applied to x / y."""
    from streammos_b200 import synthetic
    s = synthetic.make_scan(seed, n_points, t_frames)
    feats, coords, spheres = [], [], []
    for v in range(views):
        xyzi = s["xyzi"].copy()
        if v & 1:
            xyzi[..., 0] = -xyzi[..., 0]
        if v & 2:
            xyzi[..., 1] = -xyzi[..., 1]
        coord = np.stack([synthetic.quantize_bev(f) for f in xyzi])          # (T, N, 3)
        sphere = np.stack([synthetic.quantize_sphere(f) for f in xyzi])      # (T, N, 2)
        x, y, z = xyzi[..., 0], xyzi[..., 1], xyzi[..., 2]
        dist = np.sqrt(x ** 2 + y ** 2 + z ** 2) + 1e-12
        feat7 = np.stack((x, y, z, xyzi[..., 3], dist, coord[..., 0] - np.floor(coord[..., 0]),
                          coord[..., 1] - np.floor(coord[..., 1])), 1).astype(np.float32)  # (T, 7, N)
        feats.append(feat7[..., None])
        coords.append(coord[..., None])
        spheres.append(sphere[..., None])
    return {"pcds_xyzi": torch.from_numpy(np.stack(feats))[None],
            "pcds_coord": torch.from_numpy(np.stack(coords).astype(np.float32))[None],
            "pcds_sphere_coord": torch.from_numpy(np.stack(spheres).astype(np.float32))[None]}


def run_stream(net, device, batches):
    """AttNet.infer over consecutive scans with the carried query_embed_store (val_StreamMOS.py:85-95).
    -> list of (pred_cls, memory) on the CPU."""
    outs, store = [], None
    with torch.no_grad():
        for i, b in enumerate(batches):
            b = {k: v.to(device) for k, v in b.items()}
            pred_cls, _, _, _, store = net.infer(b, i, store)
            outs.append((pred_cls.float().cpu(), store.float().cpu()))
    return outs
