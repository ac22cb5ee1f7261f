"""Install the UNMODIFIED reference under baseline/_ref/ (git-ignored; travels to the GPU box with gpurun) so that
`-m gpu` tests and tools can run the reference's own model code (`models/StreamMOS.py`,
`networks/multi_view_encoder.py`, `deformattn/`) — /root/reference does not exist on the GPU box.

    python tools/install_ref.py [--ref /root/reference] [--no-cuda] [--force]

1. copies the reference tree (python sources, configs; not picture/) to baseline/_ref/StreamMOS/;
2. unless --no-cuda: compiles the reference's own CUDA extensions for sm_100a (nvcc cross-compiles without a GPU) as the
   same-box GPU baseline of tools/bench_kernels.py and tests/test_reference_model.py:
     * point_deep.cuda_kernel  <- deep_point/src/point_deep_cuda.cpp + point_deep_cuda_kernel.cu, unmodified;
     * MultiScaleDeformableAttention <- deformattn/src/**, from a temporary copy with the two `value.type()` tokens of
       ms_deform_attn_cuda.cu:64,134 replaced by `value.scalar_type()` (torch >= 2 removed the implicit conversion
       AT_DISPATCH_FLOATING_TYPES relied on). Nothing else is touched.
   Outputs go to baseline/_ref/ext/ only. The reference's own setup.py files are not run.

The pip recipe of the base contract does not apply: the reference is a script tree, not a package.
"""
import argparse
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DEST = os.path.join(ROOT, "baseline", "_ref")
TREE = os.path.join(DEST, "StreamMOS")
EXT = os.path.join(DEST, "ext")


def copy_tree(ref, force=False):
    if os.path.isdir(TREE) and not force:
        return TREE
    if os.path.isdir(TREE):
        shutil.rmtree(TREE)
    os.makedirs(DEST, exist_ok=True)
    shutil.copytree(ref, TREE, ignore=shutil.ignore_patterns("picture", ".git", "__pycache__", "*.pyc"))
    return TREE


def _load(name, sources, include_dirs, build_dir, extra_cuda=(), extra_c=()):
    os.environ.setdefault("TORCH_CUDA_ARCH_LIST", "10.0")
    from torch.utils import cpp_extension as ce
    os.makedirs(build_dir, exist_ok=True)
    ce.load(name=name, sources=sources, extra_include_paths=include_dirs, build_directory=build_dir,
            extra_cflags=["-O2", "-DVERSION_GE_1_3"] + list(extra_c),
            extra_cuda_cflags=["-O3", "-gencode", "arch=compute_100a,code=sm_100a", "-DVERSION_GE_1_3"] + list(extra_cuda),
            is_python_module=False, verbose=False)
    return os.path.join(build_dir, name + ".so")


def build_cuda(force=False):
    """-> dict name -> .so path. Compiles from the copy under baseline/_ref/StreamMOS."""
    out = {}
    dp = os.path.join(TREE, "deep_point", "src")
    so = os.path.join(EXT, "point_deep_cuda", "ref_point_deep_cuda.so")
    if force or not os.path.exists(so):
        so = _load("ref_point_deep_cuda", [os.path.join(dp, "point_deep_cuda.cpp"), os.path.join(dp, "point_deep_cuda_kernel.cu")],
                   [dp], os.path.dirname(so))
    out["point_deep_cuda"] = so
    so = os.path.join(EXT, "msda", "ref_msda.so")
    if force or not os.path.exists(so):
        src = os.path.join(TREE, "deformattn", "src")
        tmp = os.path.join(EXT, "msda", "src_patched")
        if os.path.isdir(tmp):
            shutil.rmtree(tmp)
        shutil.copytree(src, tmp)
        cu = os.path.join(tmp, "cuda", "ms_deform_attn_cuda.cu")
        text = open(cu).read()
        assert text.count("AT_DISPATCH_FLOATING_TYPES(value.type()") == 2
        open(cu, "w").write(text.replace("AT_DISPATCH_FLOATING_TYPES(value.type()", "AT_DISPATCH_FLOATING_TYPES(value.scalar_type()"))
        so = _load("ref_msda", [os.path.join(tmp, "vision.cpp"), os.path.join(tmp, "cpu", "ms_deform_attn_cpu.cpp"), cu],
                   [tmp], os.path.dirname(so),
                   extra_cuda=["-DWITH_CUDA", "-DCUDA_HAS_FP16=1", "-D__CUDA_NO_HALF_OPERATORS__",
                               "-D__CUDA_NO_HALF_CONVERSIONS__", "-D__CUDA_NO_HALF2_OPERATORS__"],
                   extra_c=["-DWITH_CUDA"])
    out["msda"] = so
    return out


def install(ref="/root/reference", cuda=True, force=False):
    if not os.path.isdir(ref):
        return None  # GPU box: whatever travelled is what there is
    copy_tree(ref, force)
    if cuda:
        build_cuda(force)
    return DEST


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", default="/root/reference")
    ap.add_argument("--no-cuda", action="store_true")
    ap.add_argument("--force", action="store_true")
    a = ap.parse_args()
    print(install(a.ref, not a.no_cuda, a.force))
