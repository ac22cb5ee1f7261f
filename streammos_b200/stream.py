"""Per-scan streaming harness for the hot path (BASELINE.json configs 2-5).

One `HotPath` object = one independent scan stream on one GPU: it owns the short-term memory (the
64x64x128 BEV feature carried between scans) and the long-term memory (ring of the last 8 scans'
points and predictions) in that GPU's HBM. Streams never talk to each other — multi-GPU runs are one
process per GPU with no collective on the data path.

The step calls the hot-path operators in the order and with the shapes of the reference inference
(models/StreamMOS.py:86-113, networks/multi_view_encoder.py:390-456, voxel_voting.py:218-243,
voxel_instance_voting.py:177-187) through the reference-shaped API of this package. The dense CNN
blocks between the operators are out of scope (cuDNN); where the reference feeds a CNN output into a
gather, a resident synthetic feature map of the same shape stands in (x0, x1, decoder output), and the
two range-view CNNs are the identity, so pool -> gather -> pool chains carry real data.
"""
import numpy as np
import torch

from . import MultiScaleDeformableAttention as MSDA
from . import deep_point, ops, synthetic, voting
from .backbone import BilinearSample

HISTORY = 8           # frames_num_max, voxel_voting.py:140
MEM_HW = 64           # query_size, multi_view_encoder.py:326
N_HEADS, HEAD_DIM, N_POINTS = 4, 32, 4
N_BOXES = 32


class ScanBatch:
    """Tensors one scan brings into the hot path (host-pinned or device)."""

    FIELDS = ("feat", "coord_bev", "coord_rv", "xyzi", "pred", "loc", "attn")

    def __init__(self, **kw):
        for f in self.FIELDS:
            setattr(self, f, kw[f])

    def nbytes(self):
        return sum(getattr(self, f).numel() * getattr(self, f).element_size() for f in self.FIELDS)

    def to(self, device, non_blocking=True):
        return ScanBatch(**{f: getattr(self, f).to(device, non_blocking=non_blocking) for f in self.FIELDS})

    def copy_from(self, other):
        for f in self.FIELDS:
            getattr(self, f).copy_(getattr(other, f), non_blocking=True)

    def empty_like(self, device):
        return ScanBatch(**{f: torch.empty_like(getattr(self, f), device=device) for f in self.FIELDS})


def make_host_scan(seed, n_points=120000, t_frames=3, channels=64, pin=True, feat_point_major=False):
    """Synthetic inputs of one scan: LiDAR-shaped coordinates, post-ReLU point features (the PointNet
    output of models/StreamMOS.py:101), predicted labels for the long-term memory and the sampling
    locations / attention weights the two deformable-attention layers receive."""
    s = synthetic.make_scan(seed, n_points, t_frames)
    rng = np.random.default_rng(seed + 7919)
    feat = np.maximum(rng.standard_normal((t_frames, channels, n_points, 1), dtype=np.float32), 0)
    coord_bev = np.ascontiguousarray(s["pcds_coord"][:, :, :2])               # (T, N, 2, 1)
    coord_rv = np.ascontiguousarray(s["pcds_sphere_coord"][:1])              # (1, N, 2, 1)
    pred = rng.integers(0, 3, n_points).astype(np.uint8)
    pred[s["n_valid"][0]:] = 0
    q = MEM_HW * MEM_HW
    ys, xs = np.meshgrid(np.linspace(0.5, MEM_HW - 0.5, MEM_HW), np.linspace(0.5, MEM_HW - 0.5, MEM_HW),
                         indexing="ij")
    ref_pts = np.stack((xs.reshape(-1) / MEM_HW, ys.reshape(-1) / MEM_HW), -1)  # multi_view_encoder.py:254-266
    loc = ref_pts[None, None, :, None, None, None, :] + \
        rng.standard_normal((2, 1, q, N_HEADS, 1, N_POINTS, 2)) * (2.0 / MEM_HW)
    a = rng.standard_normal((2, 1, q, N_HEADS, N_POINTS))
    attn = (np.exp(a) / np.exp(a).sum(-1, keepdims=True)).reshape(2, 1, q, N_HEADS, 1, N_POINTS)
    feat_t = torch.from_numpy(feat)
    if feat_point_major:  # channels_last strides: each point's channels contiguous, as the drop-in PointNet stem emits them
        feat_t = feat_t.contiguous(memory_format=torch.channels_last)
    out = ScanBatch(feat=feat_t, coord_bev=torch.from_numpy(coord_bev),
                    coord_rv=torch.from_numpy(coord_rv), xyzi=torch.from_numpy(s["xyzi"][0].copy()),
                    pred=torch.from_numpy(pred), loc=torch.from_numpy(loc.astype(np.float32)),
                    attn=torch.from_numpy(attn.astype(np.float32)))
    if pin and torch.cuda.is_available():
        out = ScanBatch(**{f: getattr(out, f).pin_memory() for f in ScanBatch.FIELDS})
    return out


class _FlatBatch:
    """Batches whose tensors can live in one flat buffer (pack()) so that a scan moves host -> device with ONE copy:
    several separate H2D copies of 0.1-10 MB leave PCIe idle between them."""

    FIELDS = ()
    _flat = None

    def __init__(self, **kw):
        for f in self.FIELDS:
            setattr(self, f, kw[f])

    def nbytes(self):
        if self._flat is not None:  # packed: the whole flat buffer (alignment padding included) is what moves
            return self._flat.numel()
        return sum(getattr(self, f).numel() * getattr(self, f).element_size() for f in self.FIELDS)

    def to(self, device, non_blocking=True):
        return type(self)(**{f: getattr(self, f).to(device, non_blocking=non_blocking) for f in self.FIELDS})

    def copy_from(self, other):
        if self._flat is not None and other._flat is not None:
            self._flat.copy_(other._flat, non_blocking=True)
            return
        for f in self.FIELDS:
            getattr(self, f).copy_(getattr(other, f), non_blocking=True)

    def pack(self, device=None, pin=False):
        """Same tensors as views of one flat byte buffer (256-byte aligned fields) on `device` (default: where they
        are), optionally pinned."""
        fields = [getattr(self, f) for f in self.FIELDS]
        offs, total = [], 0
        for t in fields:
            offs.append(total)
            total += (t.numel() * t.element_size() + 255) // 256 * 256
        dev = fields[0].device if device is None else torch.device(device)
        flat = torch.empty(total, dtype=torch.uint8, device=dev)
        if pin and dev.type == "cpu" and torch.cuda.is_available():
            flat = flat.pin_memory()
        views = {}
        for f, t, o in zip(self.FIELDS, fields, offs):
            v = flat[o:o + t.numel() * t.element_size()].view(t.dtype).view(t.shape)
            v.copy_(t)
            views[f] = v
        out = type(self)(**views)
        out._flat = flat
        return out


class LoaderBatch(_FlatBatch):
    """What the reference's val loader hands to the model for one scan (datasets/data_StreamMOS.py:565-574,
    models/StreamMOS.py:86-93): T pose-aligned frames of 7-channel point features, BEV and range-view quantised
    coordinates — plus the stand-ins for network intermediates this harness cannot produce (predicted labels,
    deformable-attention sampling locations / weights). The 64-channel point features are computed ON THE DEVICE
    by the PointNet stem (HotPath.point_pre), as in the reference: they never cross PCIe."""

    FIELDS = ("pcds_xyzi", "pcds_coord", "pcds_sphere_coord", "pred", "loc", "attn")
    coord_bev = property(lambda self: self.pcds_coord[:, :, :2])          # StreamMOS.py:102 (view, no copy)
    coord_rv = property(lambda self: self.pcds_sphere_coord[:1])          # :99


class RawBatch(_FlatBatch):
    """One scan as the loader has it BEFORE form_batch (datasets/data_StreamMOS.py:565-574): the raw, range filtered
    and padded points of the T frames, plus the range-view coordinates of the current frame (SphereQuantize stays a
    loader output: numpy's float32 arctan2 / arcsin cannot be matched bit for bit on the device) and the same
    stand-ins as LoaderBatch. Quantize + make_point_feat run on the device (ops.form_batch, bit-exact)."""

    FIELDS = ("points", "sphere_cur", "pred", "loc", "attn")
    coord_rv = property(lambda self: self.sphere_cur)


class ResidentRawBatch(_FlatBatch):
    """What has to reach the device for one scan when the stream keeps the RAW scans of its T-frame window resident in
    HBM (smos_ingest_frames): the new raw scan exactly as read from the .bin file (own sensor frame,
    no filter, no padding: `raw` (n_cap, 4) with `meta[0]` rows in use), the T pose_diff matrices of the window
    (`poses` (T, 12) float64, datasets/data_StreamMOS.py:427-447), the range-view coordinates of the current frame
    (SphereQuantize stays a loader output) and the stand-ins of RawBatch. The two older frames are the `raw` buffers of
    the previous scans (`older`, set by link_window): pose alignment, range filter and padding of all T frames run on
    the device, bit-exact with the loader (ops.ingest_frames)."""

    FIELDS = ("raw", "meta", "poses", "sphere_cur", "pred", "loc", "attn")
    coord_rv = property(lambda self: self.sphere_cur)
    older = ()


class ResidentScanBatch(_FlatBatch):
    """ResidentRawBatch without `sphere_cur`: the stream receives NOTHING the loader computed — the raw scan as read from
    the file, the window's poses and the stand-ins — and the range-view coordinates of the current frame come from
    smos_sphere_quantize on the ingested frame (HotPath(sphere_on_device=True)). Floating point, not bit-exact with
    numpy's float32 arctan2 / arcsin (include/streammos_b200.h: within 1 ulp of the angle; a handful of the 120 k
    points change their range-view cell), which is why the bit-exact ResidentRawBatch stays the default."""

    FIELDS = ("raw", "meta", "poses", "pred", "loc", "attn")
    coord_rv = None
    older = ()


def link_window(batches, t_frames=3):
    """Scan i of a cyclic list of device batches sees the raw scans of i-1, i-2, ... as the older frames of its window."""
    n = len(batches)
    for i, b in enumerate(batches):
        b.older = tuple(batches[(i - k) % n] for k in range(1, t_frames))
    return batches


def make_host_resident_stream(seed, n_scans, n_points=120000, t_frames=3, n_cap=None, pin=True, device_sphere=False):
    """`n_scans` consecutive scans of ONE synthetic drive, cyclic (scan 0 follows scan n_scans-1): the vehicle moves by
    the same rigid motion every scan, so the window's pose_diff matrices are the same for every scan. Returns
    (resident, raw): ResidentRawBatch per scan, and the RawBatch the HOST would build from the same drive today (its
    numpy pose alignment + filter + padding of all T frames) — the two describe identical model inputs.
    device_sphere=True: ResidentScanBatch (no range-view coordinates from the host)."""
    raws = [synthetic.lidar_raw_scan(np.random.default_rng(seed * 7919 + 31 * i)) for i in range(n_scans)]
    n_cap = n_cap or (max(len(r) for r in raws) + 255) // 256 * 256
    diffs = synthetic.stream_pose_diffs(t_frames)
    poses = np.stack([np.ascontiguousarray(d[:3]).reshape(12) for d in diffs])
    resident, host_aligned = [], []
    for i in range(n_scans):
        frames = [synthetic.align_filter_pad(raws[(i - k) % n_scans], diffs[k], n_points)[0] for k in range(t_frames)]
        other = make_host_scan(seed * 1000 + i, n_points, t_frames, channels=1, pin=False)
        sphere = np.ascontiguousarray(synthetic.quantize_sphere(frames[0])[None, :, :, None])     # (1, N, 2, 1)
        n_valid = int((frames[0][:, 0] > synthetic.PAD_XY).sum())
        pred = other.pred.clone()
        pred[n_valid:] = 0
        buf = np.zeros((n_cap, 4), np.float32)
        buf[: len(raws[i])] = raws[i]
        r = ResidentRawBatch(raw=torch.from_numpy(buf), meta=torch.tensor([len(raws[i]), 0, 0, 0], dtype=torch.int32),
                             poses=torch.from_numpy(poses.copy()), sphere_cur=torch.from_numpy(sphere), pred=pred,
                             loc=other.loc, attn=other.attn)
        h = RawBatch(points=torch.from_numpy(np.stack(frames)), sphere_cur=torch.from_numpy(sphere.copy()), pred=pred.clone(),
                     loc=other.loc, attn=other.attn)
        if device_sphere:
            r = ResidentScanBatch(**{f: getattr(r, f) for f in ResidentScanBatch.FIELDS})
        if pin and torch.cuda.is_available():
            r, h = r.pack(pin=True), h.pack(pin=True)
        resident.append(r)
        host_aligned.append(h)
    return resident, host_aligned


def make_host_loader_scan(seed, n_points=120000, t_frames=3, pin=True):
    """Synthetic loader output of one scan (see LoaderBatch). Point features as make_point_feat builds them
    (data_StreamMOS.py:25-50): x, y, z, intensity, dist, diff_x, diff_y."""
    s = synthetic.make_scan(seed, n_points, t_frames)
    rng = np.random.default_rng(seed + 7919)
    xyzi = s["xyzi"]                                                           # (T, N, 4)
    x, y, z = xyzi[..., 0], xyzi[..., 1], xyzi[..., 2]
    dist = np.sqrt(x ** 2 + y ** 2 + z ** 2) + 1e-12          # float32, as make_point_feat computes it
    c = s["pcds_coord"][..., 0]                                                # (T, N, 3)
    feat7 = np.stack((x, y, z, xyzi[..., 3], dist.astype(np.float32),
                      c[..., 0] - np.floor(c[..., 0]), c[..., 1] - np.floor(c[..., 1])), 1)  # (T, 7, N)
    other = make_host_scan(seed, n_points, t_frames, channels=1, pin=False)   # same pred / loc / attn stand-ins
    out = LoaderBatch(pcds_xyzi=torch.from_numpy(np.ascontiguousarray(feat7[..., None].astype(np.float32))),
                      pcds_coord=torch.from_numpy(np.ascontiguousarray(s["pcds_coord"])),
                      pcds_sphere_coord=torch.from_numpy(np.ascontiguousarray(s["pcds_sphere_coord"])),
                      pred=other.pred, loc=other.loc, attn=other.attn)
    if pin and torch.cuda.is_available():
        out = out.pack(pin=True)
    return out


def make_host_raw_scan(seed, n_points=120000, t_frames=3, pin=True):
    """Synthetic RawBatch of one scan (same scan as make_host_loader_scan(seed): the two produce identical steps)."""
    s = synthetic.make_scan(seed, n_points, t_frames)
    other = make_host_scan(seed, n_points, t_frames, channels=1, pin=False)
    out = RawBatch(points=torch.from_numpy(np.ascontiguousarray(s["xyzi"])),
                   sphere_cur=torch.from_numpy(np.ascontiguousarray(s["pcds_sphere_coord"][:1])),
                   pred=other.pred, loc=other.loc, attn=other.attn)
    if pin and torch.cuda.is_available():
        out = out.pack(pin=True)
    return out


class HotPath:
    """One scan stream. `step(batch)` runs the whole hot path for one scan on the current CUDA stream and
    returns the per-point labels after long-term voting plus the instance votes."""

    def __init__(self, device, n_points=120000, seed=0, point_major=True, vote_api="reference",
                 batch_plans=False, grids_channels_last=False, overlap_voting=False, branches=False,
                 ordered_gathers=True, ordered_rv=False, gather_taps=False, fuse_form_batch=True,
                 instance_branch=True, sphere_on_device=False, stem_sm_share=1.0):
        """batch_plans=False (default): every operator is called with the REFERENCE's arguments only
        (VoxelMaxPool(feat, ind, size, scale), BilinearSample(grid, coord)); plans are shared through the plan cache
        exactly as they are under the unmodified reference model. batch_plans=True: the explicit plan API (all five
        plans of a scan built by one batch of launches, passed as plan= / order=).
        branches=False (default): one serial chain, the reference's data flow (multi_view_encoder.py:393-417 feeds every
        stage from the previous one through its CNN blocks).
        sphere_on_device=True: raw-scan batches get their range-view coordinates from ops.sphere_quantize (SphereQuantize
        of the loader, floating point) instead of the loader's `sphere_cur`."""
        self.device = torch.device(device)
        self.sphere_on_device = sphere_on_device
        # share of the SMs the persistent PointNet-stem kernel may take (raw-scan batches). 1.0 = one CTA per SM, right
        # for a stream that runs one kernel at a time; a pipelined stream (ScanPipeline) sets ~0.6: the stem is latency
        # bound and holds its SMs against every other kernel, the HBM-bound kernels of the neighbouring scans use the rest
        self.stem_sm_share = stem_sm_share
        self.overlap_voting = overlap_voting
        self.branches = branches and self.device.type == "cuda"
        # voxel voting (voxel_voting.py) and instance voting (voxel_instance_voting.py) are two independent
        # post-processing passes over the same local map — separate scripts in the reference. They read the same staged
        # tensors and write different outputs, so the instance vote runs as a parallel branch (real data flow, unlike
        # the `branches` experiment above, which exists only because CNN outputs are stand-ins)
        self.instance_branch = (instance_branch or self.branches) and self.device.type == "cuda"
        self.ordered_gathers, self.ordered_rv = ordered_gathers, ordered_rv
        self.gather_taps = gather_taps and ordered_gathers and point_major and batch_plans
        self.fuse_form_batch = fuse_form_batch
        self._side = None
        self._branch_streams = []
        self.batch_plans = batch_plans
        self.n_points = n_points
        self.point_major = point_major
        self.vote_api = vote_api
        g = torch.Generator(device="cpu").manual_seed(seed)

        def rnd(*shape):
            return torch.randn(*shape, generator=g).relu_().to(self.device)

        # resident stand-ins for CNN activations (shapes: multi_view_encoder.py:393,408,441-448)
        self.x0 = rnd(1, 32, 256, 256)
        self.x1 = rnd(1, 64, 128, 128)
        self.dec = rnd(1, 64, 256, 256)
        if grids_channels_last:  # what a channels_last model hands to the gathers
            self.x0 = self.x0.contiguous(memory_format=torch.channels_last)
            self.x1 = self.x1.contiguous(memory_format=torch.channels_last)
            self.dec = self.dec.contiguous(memory_format=torch.channels_last)
        # PointNet stem (models/StreamMOS.py:77 PointNetStacker(7, 64, pre_bn=True, stack_num=2)), eval mode, random
        # weights and BatchNorm statistics: only used when a step starts from loader tensors. On CUDA it runs as the
        # fused smos_point_stem_forward kernel
        from .backbone import PointNetStacker
        self.stem = PointNetStacker(7, 64, pre_bn=True, stack_num=2)
        with torch.no_grad():
            for mod in self.stem.modules():
                if isinstance(mod, torch.nn.BatchNorm2d):
                    c = mod.num_features
                    mod.weight.copy_(torch.rand(c, generator=g) + 0.5)
                    mod.bias.copy_(torch.randn(c, generator=g) * 0.2)
                    mod.running_mean.copy_(torch.randn(c, generator=g) * 0.3)
                    mod.running_var.copy_(torch.rand(c, generator=g) + 0.5)
                elif isinstance(mod, torch.nn.Conv2d):
                    mod.weight.copy_(torch.randn(mod.weight.shape, generator=g) / mod.weight.shape[1] ** 0.5)
            bn0 = self.stem.layer[0].layer[0]   # the input BatchNorm sees raw metres
            bn0.running_mean.copy_(torch.tensor([0.5, -0.3, -1.2, 0.3, 18.0, 0.5, 0.5]))
            bn0.running_var.copy_(torch.tensor([300.0, 280.0, 0.8, 0.05, 150.0, 0.08, 0.08]))
        self.stem.eval().to(self.device)
        self.stem.point_major_out = point_major
        # short-term memory: previous scan's attended BEV feature, (1, 4096, 128) (mve.py:433-439)
        self.memory = torch.randn(1, MEM_HW * MEM_HW, N_HEADS * HEAD_DIM, generator=g).to(self.device)
        self.shapes = torch.tensor([[MEM_HW, MEM_HW]], dtype=torch.int64, device=self.device)
        self.lsi = torch.zeros(1, dtype=torch.int64, device=self.device)
        # long-term memory: last 8 scans (points + predictions) and the slot of the current scan
        self.local_pts = torch.empty(HISTORY + 1, n_points, 4, device=self.device)
        self.local_pred = torch.zeros(HISTORY + 1, n_points, dtype=torch.uint8, device=self.device)
        for h in range(HISTORY):
            s = synthetic.make_scan(seed * 1000 + 100 + h, n_points, 1)
            self.local_pts[h].copy_(torch.from_numpy(s["xyzi"][0]))
            r = np.random.default_rng(seed * 1000 + 200 + h).integers(0, 3, n_points).astype(np.uint8)
            self.local_pred[h].copy_(torch.from_numpy(r))
        # the "current" slot starts as a copy of the newest history scan: the first push moves it onto itself
        self.local_pts[HISTORY].copy_(self.local_pts[HISTORY - 1])
        self.local_pred[HISTORY].copy_(self.local_pred[HISTORY - 1])
        lo, hi = synthetic.synthetic_boxes(np.random.default_rng(seed + 31), N_BOXES)
        self.box_lo, self.box_hi = torch.from_numpy(lo).to(self.device), torch.from_numpy(hi).to(self.device)
        self.g_half = BilinearSample(in_dim=32, scale_rate=(0.5, 0.5))
        self.g_quarter = BilinearSample(in_dim=64, scale_rate=(0.25, 0.25))
        self.g_half.point_major_out = point_major
        self.g_quarter.point_major_out = point_major
        self.g_half.auto_order = self.g_quarter.auto_order = bool(ordered_gathers) and not batch_plans
        self.scan_index = 0
        self.iv_ws = None  # persistent accumulator of the instance votes (ops.instance_vote_workspace)
        self.size = synthetic.BEV_SHAPE
        self.mins = (synthetic.RANGE_X[0], synthetic.RANGE_Y[0], synthetic.RANGE_Z[0])
        self.deltas = tuple(float(np.float32((r[1] - r[0]) / s)) for r, s in
                            zip((synthetic.RANGE_X, synthetic.RANGE_Y, synthetic.RANGE_Z), self.size))

    # --- the three pieces of the hot path ----------------------------------------------------------
    def point_pre(self, pcds_xyzi):
        """(T, 7, N, 1) -> (T, 64, N, 1) contiguous, post-ReLU: the tensor VoxelMaxPool #1 consumes."""
        with torch.no_grad():
            return self.stem(pcds_xyzi)

    def _fork(self, n):
        """n side streams that start after everything enqueued so far on the current stream. Under CUDA-graph
        capture they become parallel branches of the graph (independent operators of one scan overlap: the
        latency-bound small kernels of one branch run under the HBM-bound kernels of another)."""
        main = torch.cuda.current_stream(self.device)
        while len(self._branch_streams) < n:
            self._branch_streams.append(torch.cuda.Stream(self.device))
        sides = self._branch_streams[:n]
        for s in sides:
            s.wait_stream(main)
        return main, sides

    def projection(self, b):
        """Cascade projection: 5 x VoxelMaxPool + 5 x BilinearSample, in the reference's order
        (models/StreamMOS.py:101-105, mve.py:393-417). In the reference the stages form ONE chain (x0 = CNN(pool #1),
        x1 = CNN(cat(x0, pool #3)), gather #5 reads the decoder output); here the CNN outputs are resident stand-ins, so
        pool #1, the two pool/gather chains and gather #5 happen to be independent: `branches=True` (an experiment,
        not the default) runs them as parallel graph branches."""
        if hasattr(b, "raw"):         # resident window: the loader's pose alignment + range filter + padding on the device
            frames = [(b.raw, b.meta[:1], b.poses[0])] + [(o.raw, o.meta[:1], b.poses[k + 1]) for k, o in enumerate(b.older)]
            b.points, b.n_valid = ops.ingest_frames(frames, synthetic.RANGE_X, synthetic.RANGE_Y, synthetic.RANGE_Z,
                                                    self.n_points, synthetic.PAD_XY, synthetic.PAD_Z)
        if hasattr(b, "points"):      # raw scan: Quantize + make_point_feat on the device, then the stem
            if self.fuse_form_batch:  # one kernel: raw points -> 64-channel features + quantised coordinates
                cap = 0
                if self.stem_sm_share < 1.0:
                    cap = max(1, int(round(self.stem_sm_share *
                                           torch.cuda.get_device_properties(self.device).multi_processor_count)))
                feat, coord = ops.point_stem_forward_raw(b.points, synthetic.RANGE_X, synthetic.RANGE_Y, synthetic.RANGE_Z,
                                                         self.size, *self.stem.fused_parameters(),
                                                         point_major_out=self.point_major, max_ctas=cap)
            else:
                feat7, coord = ops.form_batch(b.points, synthetic.RANGE_X, synthetic.RANGE_Y, synthetic.RANGE_Z, self.size)
                feat = self.point_pre(feat7)
            coord_bev = coord[:, :, :2]
        else:
            coord_bev = b.coord_bev
            feat = b.feat if hasattr(b, "feat") else self.point_pre(b.pcds_xyzi)
        cur_bev, cur_rv = coord_bev[:1], b.coord_rv
        if hasattr(b, "points") and (self.sphere_on_device or cur_rv is None):
            # SphereQuantize of the current frame (the only range-view coordinates the model reads, StreamMOS.py:99)
            cur_rv = ops.sphere_quantize(b.points[:1], theta_range=synthetic.RV_THETA, size=synthetic.RV_SHAPE)
        # all five pooling plans of the scan depend on the coordinates only: four launches build them all
        if self.batch_plans:
            pl = ops.pool_plan_multi([(coord_bev, (512, 512), (1.0, 1.0)), (cur_rv, (32, 1024), (0.5, 0.5)),
                                      (cur_bev, (256, 256), (0.5, 0.5)), (cur_rv, (16, 512), (0.25, 0.25)),
                                      (cur_bev, (128, 128), (0.25, 0.25))],
                                     gather_taps=[False] + [self.gather_taps] * 4)
        else:
            pl = [None] * 5
        # gathers visit the points in the cell order of the plan that shares their coordinates (BEV only by default:
        # range-view coordinates are already row-coherent in scan order)
        od = [p if (self.ordered_gathers and self.point_major and (i in (2, 4) or self.ordered_rv or self.gather_taps))
              else None for i, p in enumerate(pl)]

        def half_chain():
            x0_pt = self.g_half(self.x0, cur_bev, od[2])                                                 # mve.py:395
            x0_rv = deep_point.VoxelMaxPool(x0_pt, cur_rv, (32, 1024), (0.5, 0.5), pl[1])           # :396
            x0_pt2 = self.g_half(x0_rv, cur_rv, od[1])                                              # :400
            return deep_point.VoxelMaxPool(x0_pt2, cur_bev, (256, 256), (0.5, 0.5), pl[2]), (x0_pt, x0_rv, x0_pt2)

        def quarter_chain():
            x1_pt = self.g_quarter(self.x1, cur_bev, od[4])                                         # :410
            x1_rv = deep_point.VoxelMaxPool(x1_pt, cur_rv, (16, 512), (0.25, 0.25), pl[3])          # :411
            x1_pt2 = self.g_quarter(x1_rv, cur_rv, od[3])                                           # :415
            return deep_point.VoxelMaxPool(x1_pt2, cur_bev, (128, 128), (0.25, 0.25), pl[4]), x1_pt2, (x1_pt, x1_rv)

        if not self.branches:
            bev_in = deep_point.VoxelMaxPool(feat, coord_bev, (512, 512), (1.0, 1.0), pl[0])        # StreamMOS.py:102
            x0_bev, _ = half_chain()
            x1_bev, x1_pt, _ = quarter_chain()
            pt_bev = self.g_half(self.dec, cur_bev, od[2])                                          # StreamMOS.py:105
            return bev_in, x0_bev, x1_bev, x1_pt, pt_bev
        main, (s1, s2, s3) = self._fork(3)
        with torch.cuda.stream(s1):
            x0_bev, keep1 = half_chain()
        with torch.cuda.stream(s2):
            x1_bev, x1_pt, keep2 = quarter_chain()
        with torch.cuda.stream(s3):
            pt_bev = self.g_half(self.dec, cur_bev, od[2])
        bev_in = deep_point.VoxelMaxPool(feat, coord_bev, (512, 512), (1.0, 1.0), pl[0])
        for s in (s1, s2, s3):
            main.wait_stream(s)
        # intermediates stay referenced until the join so the caching allocator cannot hand their blocks out early
        self._keep = (keep1, keep2, pl, feat)
        return bev_in, x0_bev, x1_bev, x1_pt, pt_bev

    def temporal_fusion(self, b):
        """Two deformable-attention layers sampling the short-term memory (mve.py:268-273, 313-321). The second
        layer writes straight into the memory buffer (the first one has finished reading it): the attended feature
        IS the next scan's query_embed_store (mve.py:456), no copy."""
        value = self.memory.view(1, MEM_HW * MEM_HW, N_HEADS, HEAD_DIM)
        h = MSDA.ms_deform_attn_forward(value, self.shapes, self.lsi, b.loc[0], b.attn[0], 256)
        value2 = h.view(1, MEM_HW * MEM_HW, N_HEADS, HEAD_DIM)
        if self.device.type != "cuda":
            self.memory.copy_(MSDA.ms_deform_attn_forward(value2, self.shapes, self.lsi, b.loc[1], b.attn[1], 256))
            return self.memory
        return ops.ms_deform_attn_forward(value2, self.shapes, self.lsi, b.loc[1], b.attn[1], out=self.memory)

    def long_term_voting(self, b):
        """Voxel voting over 8 history scans + the current one, then per-instance votes. The script crops the local
        map to fov -/+ eps before it quantises (voxel_voting.py:225-231); the staging kernel applies the same crop
        (points outside the open box get coords -1: no vote, no label — also what happens to the padded points)."""
        cur = HISTORY
        # raw points of the current frame: loader batches carry them as the first 4 channels of pcds_xyzi
        if hasattr(b, "points"):
            xyzi = b.points[0]
        else:
            xyzi = b.xyzi if hasattr(b, "xyzi") else b.pcds_xyzi[0, :4, :, 0].t().contiguous()
        n = self.n_points
        pts = self.local_pts.view(-1, 4)
        prev = (self.scan_index - 1) % HISTORY
        if self.device.type == "cuda" and self.vote_api == "reference":
            # voxel_voting.py:234-243 through the reference's function signatures. One kernel inserts the new scan into
            # the ring (the previous scan moves from the current slot into its history slot, exactly as :182 walks the
            # window) and produces the script's two int64 casts (:240-241); Quantize's float32 result is a temporary
            # that dies at the cast, so it is not materialised (quantize_staged(want_q=True) returns it: tested)
            q, coords, labels = voting.quantize_staged(self.local_pts, self.local_pred, synthetic.RANGE_X,
                                                       synthetic.RANGE_Y, synthetic.RANGE_Z, self.size, new_points=xyzi,
                                                       new_pred=b.pred, cur_slot=cur, hist_slot=prev, want_q=False,
                                                       crop_eps=1e-4)
            if self.iv_ws is None:
                self.iv_ws = ops.instance_vote_workspace(N_BOXES, self.device)
            if self.instance_branch:  # the instance votes do not depend on the voxel votes: a parallel branch
                main, (side,) = self._fork(1)
                with torch.cuda.stream(side):
                    sums = ops.instance_vote(pts, labels, self.box_lo, self.box_hi, workspace=self.iv_ws)
            vl = voting.determine_voxel_labels(coords, labels, self.size, num_classes=3)
            point_labels = voting.get_point_labels_from_voxel_labels(coords[cur * n:], vl, self.size)
            if self.instance_branch:
                main.wait_stream(side)
                sums.record_stream(main)  # allocated on the side stream, consumed (D2H) on the main one
            else:
                sums = ops.instance_vote(pts, labels, self.box_lo, self.box_hi, workspace=self.iv_ws)
            return point_labels, sums
        if self.device.type == "cuda":
            ops.memory_push(xyzi, b.pred, self.local_pts[cur], self.local_pred[cur], self.local_pts[prev],
                            self.local_pred[prev])
        else:
            self.local_pts[cur].copy_(xyzi)
            self.local_pred[cur].copy_(b.pred)
        labels = self.local_pred.view(-1).to(torch.int64)              # voxel_voting.py:241
        if self.vote_api == "reference":  # CPU harness state (tests)
            q = voting.Quantize(pts, synthetic.RANGE_X, synthetic.RANGE_Y, synthetic.RANGE_Z, self.size)
            coords = q.to(torch.int64)                                 # voxel_voting.py:240
            vl = voting.determine_voxel_labels(coords, labels, self.size, num_classes=3)
            point_labels = voting.get_point_labels_from_voxel_labels(coords[cur * n:], vl, self.size)
        else:  # fused streaming variant: float xyz + uint8 labels in, no int64 staging
            _, point_labels = ops.vote_fused(pts, self.local_pred.view(-1), n, self.mins, self.deltas, self.size, 3)
        sums = ops.instance_vote(pts, labels, self.box_lo, self.box_hi)
        if self.device.type != "cuda":  # CPU harness state (tests): the current scan becomes history at the end
            slot = self.scan_index % HISTORY
            self.local_pts[slot].copy_(self.local_pts[cur])
            self.local_pred[slot].copy_(self.local_pred[cur])
        return point_labels, sums

    def step(self, b):
        """One scan. With `overlap_voting` the long-term voting runs on a second stream next to the
        projection + temporal fusion: voting post-processes PREDICTIONS (an input here, the network's argmax
        in the reference), so in a stream it is the voting of scan t-1 that overlaps the network of scan t —
        the two branches share no data. Under CUDA-graph capture the fork/join becomes two graph branches."""
        if self.overlap_voting and self.device.type == "cuda":
            main = torch.cuda.current_stream(self.device)
            if self._side is None:
                self._side = torch.cuda.Stream(self.device)
            self._side.wait_stream(main)
            with torch.cuda.stream(self._side):
                point_labels, sums = self.long_term_voting(b)
            proj = self.projection(b)
            self.temporal_fusion(b)
            main.wait_stream(self._side)
        else:
            proj = self.projection(b)
            self.temporal_fusion(b)  # the attended feature becomes the next scan's query_embed_store (mve.py:456)
            point_labels, sums = self.long_term_voting(b)
        self.scan_index += 1
        return point_labels, sums, proj


def algorithmic_bytes(n_points, t_frames=3):
    """Compulsory HBM bytes of one scan (SURVEY §8d formulas, fp32, read-once / write-once)."""
    N = n_points
    pools = [(t_frames, 64, 512, 512), (1, 32, 32, 1024), (1, 32, 256, 256), (1, 64, 16, 512), (1, 64, 128, 128)]
    gathers = [(32, 256, 256), (32, 32, 1024), (64, 128, 128), (64, 16, 512), (64, 256, 256)]
    pool = [4 * B * C * N + 8 * B * N + 4 * B * C * H * W for (B, C, H, W) in pools]
    gather = [4 * C * H * W + 8 * N + 4 * C * N for (C, H, W) in gathers]
    S = Q = MEM_HW * MEM_HW
    msda = 2 * 4 * (S * N_HEADS * HEAD_DIM + 3 * Q * N_HEADS * N_POINTS + Q * N_HEADS * HEAD_DIM)
    P = (HISTORY + 1) * N
    vote = 24 * P + 8 * P + 8 * 512 * 512 * 30 + 40 * N
    return dict(pool=pool, gather=gather, msda=msda, vote=vote,
                total=sum(pool) + sum(gather) + msda + vote)
