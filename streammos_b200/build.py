"""In-tree build of libstreammos_b200.so (hand-written sm_100a kernels + C-ABI).

    python -m streammos_b200.build            # build if stale
    python -m streammos_b200.build --force

nvcc cross-compiles without a GPU; the .so is git-ignored but travels to the GPU box.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB_DIR = os.path.join(HERE, "lib")
LIB_PATH = os.path.join(LIB_DIR, "libstreammos_b200.so")
SOURCES = ["api.cu", "voxel_maxpool.cu", "bilinear_gather.cu", "ms_deform_attn.cu", "voting.cu", "point_stem.cu", "form_batch.cu",
           "cluster.cu", "ingest.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC",
    "-Xptxas", "-v",
]


def _nvcc():
    cand = os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "bin", "nvcc")
    return cand if os.path.exists(cand) else "nvcc"


def _stale():
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)]
    deps.append(os.path.join(HERE, "..", "include", "streammos_b200.h"))
    deps.append(os.path.abspath(__file__))
    return any(os.path.getmtime(d) > t for d in deps)


def build_variant(name, defines):
    """Experiment builds for A/B runs inside one GPU call (SMOS_LIB=...): the library compiled with extra -D flags into
    lib/var_<name>/libstreammos_b200.so."""
    global LIB_DIR, LIB_PATH, NVCC_FLAGS
    saved = (LIB_DIR, LIB_PATH, NVCC_FLAGS)
    try:
        LIB_DIR = os.path.join(saved[0], "var_" + name)
        LIB_PATH = os.path.join(LIB_DIR, "libstreammos_b200.so")
        NVCC_FLAGS = saved[2] + list(defines)
        return build(force=True)
    finally:
        LIB_DIR, LIB_PATH, NVCC_FLAGS = saved


def build(force=False, verbose=False):
    if not force and not _stale():
        return LIB_PATH
    os.makedirs(LIB_DIR, exist_ok=True)
    from concurrent.futures import ThreadPoolExecutor
    # headers every translation unit may include: a change there recompiles everything, a change in one .cu only it
    shared = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if not f.endswith(".cu")]
    shared += [os.path.join(HERE, "..", "include", "streammos_b200.h"), os.path.abspath(__file__)]
    t_shared = max(os.path.getmtime(d) for d in shared)

    def compile_one(src):
        obj = os.path.join(LIB_DIR, src.replace(".cu", ".o"))
        log = obj + ".log"
        path = os.path.join(CSRC, src)
        if not force and os.path.exists(obj) and os.path.exists(log) and \
                os.path.getmtime(obj) > max(t_shared, os.path.getmtime(path)):
            return obj, open(log).read()
        cmd = [_nvcc()] + NVCC_FLAGS + ["-c", path, "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError("nvcc failed on %s" % src)
        with open(log, "w") as f:
            f.write(r.stderr)
        return obj, r.stderr

    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        done = list(ex.map(compile_one, SOURCES))
    objs = [o for o, _ in done]
    logs = [l for _, l in done]
    cmd = [_nvcc(), "-shared", "-o", LIB_PATH] + objs + ["-gencode", "arch=compute_100a,code=sm_100a"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("link failed")
    with open(os.path.join(LIB_DIR, "ptxas.log"), "w") as f:
        f.write("\n".join(logs))
    if verbose:
        sys.stderr.write("\n".join(logs))
    return LIB_PATH


if __name__ == "__main__":
    if "--variant" in sys.argv:  # python -m streammos_b200.build --variant <name> -DX=1 -DY=2
        i = sys.argv.index("--variant")
        print(build_variant(sys.argv[i + 1], sys.argv[i + 2:]))
    else:
        print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
