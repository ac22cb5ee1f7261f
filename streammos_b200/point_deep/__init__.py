"""Drop-in for the reference's compiled `point_deep` package (deep_point/setup.py:12-16):
`point_deep.cuda_kernel` is served by the sm_100a library, `point_deep.cpu_kernel` raises."""
from . import cuda_kernel, cpu_kernel  # noqa: F401
