"""The C-ABI library loads and exports every symbol include/streammos_b200.h declares, with the
parameter counts the ctypes binding uses. No compute calls (no GPU needed)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "streammos_b200.h")


def header_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    src = re.sub(r"#.*", "", src)
    out = {}
    for m in re.finditer(r"\b(smos_\w+)\s*\(([^;{]*?)\)\s*;", src, flags=re.S):
        args = m.group(2).strip()
        n = 0 if args in ("", "void") else len([a for a in args.split(",") if a.strip()])
        out[m.group(1)] = n
    return out


@pytest.fixture(scope="module")
def lib():
    from streammos_b200 import build, _lib
    build.build()
    return _lib.load()


def test_header_declares_the_hot_path():
    fns = header_functions()
    for required in ("smos_pool_plan_build", "smos_voxel_maxpool_forward", "smos_voxel_maxpool_backward",
                     "smos_bilinear_gather_forward", "smos_bilinear_gather_backward",
                     "smos_ms_deform_attn_forward", "smos_ms_deform_attn_backward", "smos_vote_voxel_labels",
                     "smos_vote_point_labels", "smos_instance_vote", "smos_quantize"):
        assert required in fns


def test_every_declared_symbol_is_exported_and_bound(lib):
    from streammos_b200 import _lib
    fns = header_functions()
    assert len(fns) >= 15
    for name, nargs in fns.items():
        assert hasattr(lib, name), "not exported: " + name
        assert name in _lib.SIGNATURES, "no ctypes signature: " + name
        assert len(_lib.SIGNATURES[name][1]) == nargs, "argument count mismatch: " + name
    assert set(_lib.SIGNATURES) == set(fns)


def test_host_only_entry_points(lib):
    assert lib.smos_abi_version() == 2
    assert lib.smos_error_string(0) == b"ok"
    assert b"invalid" in lib.smos_error_string(-1)
    assert lib.smos_pool_plan_bytes(3, 160000, 512, 512) > 3 * 160000 * 16
    assert lib.smos_pool_plan_bytes(0, 10, 4, 4) == -1
    assert lib.smos_pool_workspace_bytes(3, 64, 120000) >= 3 * 64 * 120000 * 4
    assert lib.smos_pool_workspace_bytes(3, 0, 120000) == -1
    assert lib.smos_vote_workspace_bytes(1080000, 512, 512, 30, 3) == 512 * 512 * 30 * 8
    assert lib.smos_vote_workspace_bytes(1080000, 512, 512, 30, 5) == 512 * 512 * 30 * 5 * 4
    assert lib.smos_vote_workspace_bytes(10, 0, 1, 1, 3) == -1


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    from streammos_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(RuntimeError, match="no CPU/PyTorch fallback"):
        _lib.load()
