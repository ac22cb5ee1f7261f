"""`point_deep.cpu_kernel` stand-in: the B200 build has no CPU path (north_star: no CPU fallback)."""


def _no_cpu(*_a, **_k):
    raise RuntimeError("point_deep.cpu_kernel is not available in streammos_b200: the hot path is CUDA-only "
                       "(the CPU restatement is test infrastructure under oracle/)")


voxel_maxpooling_cpu_forward = _no_cpu
voxel_maxpooling_cpu_backward = _no_cpu
