"""Pooling-plan cache: lets UNMODIFIED reference code share plans between its pooling and gather calls.

The reference calls `deep_point.VoxelMaxPool(feat, ind, size, scale)` and `BilinearSample(grid, coord)` with no notion of
a plan (models/StreamMOS.py:101-105, networks/multi_view_encoder.py:393-417), but inside one scan it passes the SAME
coordinate tensors again and again: `pcds_cood_cur` to pools #3/#5 and gathers #1/#3/#5, `pcds_sphere_coord_cur` to pools
#2/#4 and gathers #2/#4. A plan (cell index + counting sort of the points, ops.PoolPlan) depends only on
(coordinates, grid size, scale), so it is built once per such triple and found again by the later calls:

  * key   = (device, data_ptr, shape, strides, H, W, scale_h, scale_w, capture id);
  * valid = the cached entry's coordinate tensor still has the same `_version` (torch bumps it on every in-place write);
    the entry keeps a strong reference to that tensor, so its memory cannot be freed and handed to a different tensor
    while the entry lives — equal pointer + equal version therefore means equal contents;
  * CUDA graphs: an entry remembers the capture it was created in (smos_stream_capture_id; 0 = eager) and is only ever
    returned to calls of the same capture. A plan built eagerly is never baked into a graph whose replays will see new
    coordinates at the same address, and a plan built inside a capture (its kernels are part of the graph, so replays
    rebuild it) is never used by eager code.

Prefetch. Building a plan is four short dependent kernels; several plans built in one batch cost the same four
launches (ops.pool_plan_multi). The cache therefore remembers, per "kind" of coordinate tensor — (shape, strides,
geometry of the FIRST plan requested on it) — which other geometries were requested on the same tensor afterwards, and
the next time a tensor of that kind shows up (the next scan) it builds all of them with one batch of launches. A
prefetched plan that is never used is forgotten again. No model knowledge is involved: the pattern is learned from the
calls (for StreamMOS: {BEV 1/2, BEV 1/4} on pcds_cood_cur, {RV 1/2, RV 1/4} on pcds_sphere_coord_cur).

Nothing here synchronises or reads device memory. Inference-mode tensors (no version counter) are not cached.
"""
import collections
import ctypes

import torch

from . import _lib

CAPACITY = 48  # entries; a scan needs five plans, four scans are in flight, captures leave theirs behind until evicted

_entries = collections.OrderedDict()   # key -> [coordinate view, version, plan, kind, prefetched-and-unused]
_history = {}                          # kind -> ordered list of geometries (H, W, sh, sw) seen on tensors of that kind
_stats = {"hits": 0, "misses": 0, "builds": 0, "prefetched": 0}
enabled = True
prefetch = True
MAX_BATCH = 8                          # smos_pool_plan_build_multi builds at most 8 plans per call


def capture_id():
    """0 when the current stream is not capturing, else the unique id of the capture sequence."""
    out = ctypes.c_uint64(0)
    rc = _lib.load().smos_stream_capture_id(ctypes.c_void_p(torch.cuda.current_stream().cuda_stream), ctypes.byref(out))
    _lib.check(rc, "smos_stream_capture_id")
    return int(out.value)


def _version(t):
    try:
        return t._version
    except RuntimeError:  # inference tensors do not track versions
        return None


def _key(ind, H, W, scale, cap):
    return (ind.device.index, ind.data_ptr(), tuple(ind.shape), tuple(ind.stride()), int(H), int(W), float(scale[0]),
            float(scale[1]), cap)


def clear():
    _entries.clear()
    _history.clear()


def stats():
    return dict(_stats, entries=len(_entries))


def lookup(ind, output_size, scale_rate, cap=None):
    """The cached plan for exactly this (coordinates, grid, scale), or None."""
    if not enabled:
        return None
    v = _version(ind)
    if v is None:
        return None
    cap = capture_id() if cap is None else cap
    k = _key(ind, output_size[0], output_size[1], scale_rate, cap)
    e = _entries.get(k)
    if e is None:
        return None
    tensor, version, plan = e[0], e[1], e[2]
    if version != v or tensor.data_ptr() != ind.data_ptr():
        del _entries[k]
        return None
    _entries.move_to_end(k)
    e[4] = False
    _stats["hits"] += 1
    return plan


def _evict(k):
    e = _entries.pop(k)
    if e[4] and e[3] in _history:  # prefetched and never used: stop prefetching this geometry for this kind
        geo = k[4:8]
        if geo in _history[e[3]]:
            _history[e[3]].remove(geo)


def store(ind, output_size, scale_rate, plan, cap=None, kind=None, prefetched=False):
    if not enabled:
        return
    v = _version(ind)
    if v is None:
        return
    cap = capture_id() if cap is None else cap
    # plans of captures that are over can never be used again: drop them first
    for k in [k for k in _entries if k[-1] not in (0, cap)]:
        _evict(k)
    _entries[_key(ind, output_size[0], output_size[1], scale_rate, cap)] = [ind, v, plan, kind, prefetched]
    while len(_entries) > CAPACITY:
        _evict(next(iter(_entries)))


def _kind_of(ind, cap):
    """Kind of a coordinate tensor that already has live entries (same memory, same version, same capture), or None."""
    v = _version(ind)
    head = (ind.device.index, ind.data_ptr(), tuple(ind.shape), tuple(ind.stride()))
    for k, e in _entries.items():
        if k[:4] == head and k[-1] == cap and e[1] == v:
            return e[3]
    return None


def get(ind, output_size, scale_rate, build, build_multi=None):
    """Cached plan or a freshly built one (stored). `ind` is the (B, N, 2) coordinate view the plan kernels read;
    `build()` builds this one plan, `build_multi(geometries)` several plans of `ind` with one batch of launches."""
    if not enabled:
        return build()
    cap = capture_id()
    plan = lookup(ind, output_size, scale_rate, cap)
    if plan is not None:
        return plan
    _stats["misses"] += 1
    if _version(ind) is None:
        return build()
    geo = (int(output_size[0]), int(output_size[1]), float(scale_rate[0]), float(scale_rate[1]))
    kind = _kind_of(ind, cap)
    if kind is not None:  # a tensor we know already, asked for a geometry nobody predicted: learn it
        if geo not in _history.setdefault(kind, []):
            _history[kind].append(geo)
        plan = build()
        _stats["builds"] += 1
        store(ind, output_size, scale_rate, plan, cap, kind)
        return plan
    # first request on this tensor: its kind is (shape, strides, this geometry)
    kind = (tuple(ind.shape), tuple(ind.stride()), geo)
    others = [g for g in _history.setdefault(kind, []) if g != geo][:MAX_BATCH - 1] if prefetch else []
    if not others or build_multi is None:
        plan = build()
        _stats["builds"] += 1
        store(ind, output_size, scale_rate, plan, cap, kind)
        return plan
    plans = build_multi([geo] + others)
    _stats["builds"] += 1
    _stats["prefetched"] += len(others)
    for g, p in zip([geo] + others, plans):
        store(ind, (g[0], g[1]), (g[2], g[3]), p, cap, kind, prefetched=g != geo)
    return plans[0]
