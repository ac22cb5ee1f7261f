"""Compile the reference's own CPU VoxelMaxPool (deep_point/src/point_deep.cpp) UNMODIFIED, straight
from the read-only reference tree, into oracle/_ref/point_deep_cpu_ref*.so with plain g++.

Test infrastructure only: it validates the C restatement (oracle/smos_oracle.c) and can serve as the
`cpu_baseline.kind == "reference"` leg of bench.py. No reference source is copied into this repo.
The reference's own build system (setup.py) is not run.
"""
import argparse
import os
import subprocess
import sys
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
OUT_DIR = os.path.join(HERE, "_ref")
MOD_NAME = "point_deep_cpu_ref"


def out_path():
    return os.path.join(OUT_DIR, MOD_NAME + sysconfig.get_config_var("EXT_SUFFIX"))


def build(ref_root="/root/reference", force=False):
    src = os.path.join(ref_root, "deep_point", "src", "point_deep.cpp")
    if not os.path.exists(src):
        return None  # reference tree absent (GPU box): use the prebuilt file if it travelled
    out = out_path()
    if os.path.exists(out) and not force and os.path.getmtime(out) > os.path.getmtime(src):
        return out
    import torch
    from torch.utils import cpp_extension as ce

    os.makedirs(OUT_DIR, exist_ok=True)
    inc = ce.include_paths() + [sysconfig.get_paths()["include"]]
    libdir = os.path.join(os.path.dirname(torch.__file__), "lib")
    cmd = ["g++", "-O2", "-fPIC", "-shared", "-std=c++17", "-DVERSION_GE_1_3",
           "-DTORCH_EXTENSION_NAME=" + MOD_NAME, "-DTORCH_API_INCLUDE_EXTENSION_H",
           "-D_GLIBCXX_USE_CXX11_ABI=%d" % int(torch._C._GLIBCXX_USE_CXX11_ABI)]
    for i in inc:
        cmd += ["-isystem", i]
    cmd += [src, "-o", out, "-L" + libdir, "-Wl,-rpath," + libdir,
            "-lc10", "-ltorch_cpu", "-ltorch", "-ltorch_python"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stderr[-4000:])
        raise RuntimeError("building the reference point_deep.cpp failed")
    return out


def load():
    """Import the compiled reference module (None if it was never built)."""
    p = out_path()
    if not os.path.exists(p):
        return None
    import importlib.util
    import torch  # noqa: F401  (libtorch must be loaded first)
    spec = importlib.util.spec_from_file_location(MOD_NAME, p)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", default="/root/reference")
    ap.add_argument("--force", action="store_true")
    a = ap.parse_args()
    print(build(a.ref, a.force))
