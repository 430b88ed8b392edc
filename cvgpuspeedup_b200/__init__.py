"""cvgpuspeedup_b200 -- B200-native fused image preprocessing behind the cvGS operator interface.

The product is ``libcvgs_b200.so`` (hand-written sm_100a kernels + C-ABI, ``include/cvgs_b200.h``);
this package is the thin host-side mirror of ``namespace cvGS`` used by the tests and the benchmark.
"""
from . import _abi  # noqa: F401
from .api import *  # noqa: F401,F403
from .api import (CircularTensor, GpuMat, add, build_pipeline, convertTo, cvtColor, divide,  # noqa: F401
                  executeOperations, make_crops, multiply, resize, resize_nv12, split, split_planes, splitT, subtract,
                  warp, write, write_u8)
