"""jrender_b200 -- B200-native differentiable mesh rasterizer behind Jittor/jrender's API.

Hot path: hand-written sm_100a CUDA kernels in csrc/, exposed through the C ABI in
include/b200raster.h (libb200raster.so) and bound with ctypes; PyTorch is plumbing only.
"""
from .softras import SoftRasterizeFunction, SoftRasterizer, soft_rasterize  # noqa: F401
from ._lib import B200RasterError  # noqa: F401
