"""ctypes binding of libstreammos_b200.so (include/streammos_b200.h).

The library is the ONLY compute path: if it is missing, loading raises — there is no CPU or
PyTorch fallback (BASELINE.json north_star: "no CPU fallback").
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libstreammos_b200.so")
LIB_PATH = os.environ.get("SMOS_LIB", LIB_PATH)  # experiments: A/B two builds in one GPU call

_i32, _i64, _f32, _vp = ctypes.c_int32, ctypes.c_int64, ctypes.c_float, ctypes.c_void_p

class PoolPlanDesc(ctypes.Structure):
    """struct smos_pool_plan_desc (include/streammos_b200.h)."""
    _fields_ = [("pcds_ind", _vp), ("B", _i64), ("N", _i64), ("ind_sb", _i64), ("ind_sn", _i64), ("ind_sd", _i64),
                ("H", _i32), ("W", _i32), ("scale_h", _f32), ("scale_w", _f32), ("voxel_max_idx", _vp),
                ("idx_batch_stride", _i64), ("plan", _vp), ("gather_taps", _vp), ("scale_dev", _vp)]


class VoteStreamScan(ctypes.Structure):
    """struct smos_vote_stream_scan (include/streammos_b200.h)."""
    _fields_ = [("points", _vp), ("labels", _vp), ("n", _i64), ("pose_diff", ctypes.c_double * 12),
                ("transform", _i32)]


class IngestFrame(ctypes.Structure):
    """struct smos_ingest_frame (include/streammos_b200.h)."""
    _fields_ = [("points", _vp), ("n_dev", _vp), ("pose_dev", _vp), ("n_cap", _i64)]


# name -> (restype, argtypes); mirrors include/streammos_b200.h one to one
SIGNATURES = {
    "smos_abi_version": (ctypes.c_int, []),
    "smos_error_string": (ctypes.c_char_p, [ctypes.c_int]),
    "smos_stream_capture_id": (ctypes.c_int, [_vp, ctypes.POINTER(ctypes.c_uint64)]),
    "smos_pool_plan_bytes": (_i64, [_i64, _i64, _i32, _i32]),
    "smos_pool_workspace_bytes": (_i64, [_i64, _i64, _i64]),
    "smos_pool_plan_build": (ctypes.c_int, [_vp, _i64, _i64, _i64, _i64, _i64, _i32, _i32, _f32, _f32, _vp, _i64,
                                            _vp, _vp]),
    "smos_pool_plan_build_multi": (ctypes.c_int, [ctypes.POINTER(PoolPlanDesc), _i32, _vp]),
    "smos_voxel_maxpool_forward": (ctypes.c_int, [_vp, _i64, _i64, _i64, _i64, _i64, _i64, _i32, _i32, _vp, _vp,
                                                  _vp, _vp]),
    "smos_voxel_maxpool_forward_stages": (ctypes.c_int, [_vp, _i64, _i64, _i64, _i64, _i64, _i64, _i32, _i32, _vp,
                                                         _vp, _vp, _i32, _vp]),
    "smos_voxel_maxpool_backward": (ctypes.c_int, [_vp, _i64, _i64, _i64, _i64, _i64, _i64, _i32, _i32, _vp, _vp,
                                                   _vp, _vp, _i64, _i64, _i64, _vp]),
    "smos_bilinear_gather_forward": (ctypes.c_int, [_vp, _i64, _i64, _i32, _i32, _i64, _i64, _i64, _i64, _vp, _i64,
                                                    _i64, _i64, _i64, _f32, _f32, _vp, _i64, _i64, _i64, _vp]),
    "smos_bilinear_gather_forward_ordered": (ctypes.c_int, [_vp, _i64, _i64, _i32, _i32, _i64, _i64, _i64, _i64, _vp, _i64,
                                                    _i64, _i64, _i64, _f32, _f32, _vp, _i64, _i64, _i64, _vp, _i32, _i32, _vp]),
    "smos_gather_taps_bytes": (ctypes.c_int64, [_i64, _i64]),
    "smos_bilinear_gather_forward_taps": (ctypes.c_int, [_vp, _i64, _i64, _i32, _i32, _i64, _i64, _i64, _i64, _vp, _i64,
                                                         _vp, _i64, _i64, _i64, _vp]),
    "smos_bilinear_gather_backward": (ctypes.c_int, [_vp, _i64, _i64, _i64, _i64, _i64, _i64, _vp, _i64, _i64, _i64,
                                                     _f32, _f32, _i32, _i32, _vp, _vp]),
    "smos_ms_deform_attn_forward": (ctypes.c_int, [_i32, _vp, _vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _i32,
                                                   _i32, _i32, _vp, _vp]),
    "smos_ms_deform_attn_fused_forward": (ctypes.c_int, [_i32, _vp, _vp, _vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _i32,
                                                         _i32, _i32, _i32, _vp, _vp]),
    "smos_ms_deform_attn_backward": (ctypes.c_int, [_i32, _vp, _vp, _vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32,
                                                    _i32, _i32, _i32, _vp, _vp, _vp, _vp]),
    "smos_quantize": (ctypes.c_int, [_vp, _i64, _i64, _f32, _f32, _f32, _f32, _f32, _f32, _vp, _vp]),
    "smos_quantize_rcp": (ctypes.c_int, [_vp, _i64, _i64, _f32, _f32, _f32, _f32, _f32, _f32, _vp, _vp]),
    "smos_vote_workspace_bytes": (_i64, [_i64, _i32, _i32, _i32, _i32]),
    "smos_vote_voxel_labels": (ctypes.c_int, [_vp, _vp, _i64, _i32, _i32, _i32, _i32, _vp, _vp, _vp]),
    "smos_vote_point_labels": (ctypes.c_int, [_vp, _i64, _vp, _i32, _i32, _i32, _vp, _vp]),
    "smos_vote_fused": (ctypes.c_int, [_vp, _i64, _i64, _vp, _i64, _f32, _f32, _f32, _f32, _f32, _f32, _i32, _i32,
                                       _i32, _i32, _vp, _vp, _vp, _vp]),
    "smos_vote_stream": (ctypes.c_int, [ctypes.POINTER(VoteStreamScan), _i32, _i32, _i64, ctypes.POINTER(_f32),
                                        ctypes.POINTER(_f32), _f32, _f32, _f32, _f32, _f32, _f32, _i32, _i32, _i32, _i32,
                                        _vp, _vp, _vp, _vp]),
    "smos_point_stem_forward": (ctypes.c_int, [_vp, _i64, _i32, _i64, _i64, _i64, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp,
                                               _vp, _i32, _i32, _vp, _i64, _i64, _i64, _vp]),
    "smos_form_batch": (ctypes.c_int, [_vp, _i64, _i64, _i64, _f32, _f32, _f32, _f32, _f32, _f32, _f32, _f32, _vp, _vp, _vp]),
    "smos_sphere_quantize": (ctypes.c_int, [_vp, _i64, _i64, _f32, _f32, _f32, _f32, _f32, _f32, _vp, _vp]),
    "smos_ingest_workspace_bytes": (_i64, [_i32, _i64, _i64]),
    "smos_ingest_frames": (ctypes.c_int, [ctypes.POINTER(IngestFrame), _i32, _i64, _f32, _f32, _f32, _f32, _f32, _f32, _i64,
                                          _f32, _f32, _vp, _vp, _vp, _vp, _vp]),
    "smos_point_stem_forward_raw": (ctypes.c_int, [_vp, _i64, _i64, _i64, _f32, _f32, _f32, _f32, _f32, _f32, _f32, _f32,
                                                   _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i32, _i32, _vp, _vp, _i64, _i64,
                                                   _i64, _vp]),
    "smos_point_stem_forward_raw_capped": (ctypes.c_int, [_vp, _i64, _i64, _i64, _f32, _f32, _f32, _f32, _f32, _f32, _f32,
                                                          _f32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i32, _i32, _vp, _vp,
                                                          _i64, _i64, _i64, _i32, _vp]),
    "smos_vote_stage": (ctypes.c_int, [_vp, _vp, _i32, _i64, _i64, _vp, _vp, _i32, _i32, _f32, _f32, _f32, _f32, _f32, _f32,
                                       ctypes.POINTER(_f32), ctypes.POINTER(_f32), _vp, _vp, _vp, _vp]),
    "smos_memory_push": (ctypes.c_int, [_vp, _vp, _i64, _i64, _vp, _vp, _vp, _vp, _vp]),
    "smos_instance_vote": (ctypes.c_int, [_vp, _i64, _i64, _vp, _vp, _vp, _i32, _vp, _vp]),
    "smos_instance_vote_counted": (ctypes.c_int, [_vp, _i64, _i64, _vp, _vp, _vp, _i32, _vp, _vp, _vp]),
    "smos_instance_vote_workspace_bytes": (_i64, [_i32]),
    "smos_instance_vote_ws": (ctypes.c_int, [_vp, _i64, _i64, _vp, _vp, _vp, _i32, _vp, _vp, _vp, _vp]),
    "smos_cluster_workspace_bytes": (ctypes.c_int64, [_i64]),
    "smos_cluster_boxes": (ctypes.c_int, [_vp, _i64, _i64, _vp, ctypes.c_double, _i32, _i32, ctypes.c_float, _vp,
                                          _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "smos_cluster_apply": (ctypes.c_int, [_i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
}

ABI_VERSION = 2  # SMOS_ABI_VERSION of include/streammos_b200.h
_lib = None


def load():
    """Load the shared library (once). Raises RuntimeError if it was not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            "streammos_b200: %s not found — build it with `python -m streammos_b200.build` "
            "(there is no CPU/PyTorch fallback for the hot path)" % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the header and the .so disagree
        fn.restype = res
        fn.argtypes = args
    if lib.smos_abi_version() != ABI_VERSION:
        raise RuntimeError("streammos_b200: ABI version mismatch")
    _lib = lib
    return lib


def check(code, what):
    if code != 0:
        msg = load().smos_error_string(int(code))
        raise RuntimeError("%s failed: %s (code %d)" % (what, msg.decode() if msg else "?", code))
