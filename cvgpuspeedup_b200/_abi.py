"""ctypes view of include/cvgs_b200.h (the C-ABI of libcvgs_b200.so).

The library is the product; there is no CPU fallback.  Importing this module
fails loudly when the shared library has not been built (run
``python -c "import __graft_entry__ as g; g.build()"`` or ``make -C cvgpuspeedup_b200``).
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("CVGS_B200_LIB") or os.path.join(_HERE, "libcvgs_b200.so")  # override: diagnostic builds

MAX_OPS = 8

# enums (include/cvgs_b200.h)
CVGS_8UC3, CVGS_16UC3, CVGS_16SC3, CVGS_32FC3 = 16, 18, 19, 21
CVGS_8UC4, CVGS_16UC4, CVGS_16SC4, CVGS_32FC4 = 24, 26, 27, 29
CVGS_NV12 = 0x1001
CVGS_NV21, CVGS_P010, CVGS_P210, CVGS_Y210 = 0x1002, 0x1003, 0x1004, 0x1005
YUV_BT601_FULL, YUV_BT709_FULL, YUV_BT709_LIMITED, YUV_BT2020_FULL = 0, 1, 2, 3
WARP_AFFINE, WARP_PERSPECTIVE = 0, 1
PRESERVE_AR, IGNORE_AR, PRESERVE_AR_RN_EVEN, PRESERVE_AR_LEFT = 0, 1, 2, 3
OP_MUL, OP_SUB, OP_DIV, OP_ADD, OP_REORDER = 1, 2, 3, 4, 5
OP_ADD_ALPHA, OP_DROP_ALPHA, OP_GRAY = 6, 7, 8
CVGS_32FC1 = 5
FP_REFERENCE_FUSED, FP_SEPARATE = 0, 1
INTERP_FLOAT, INTERP_ROUND_U8 = 0, 1
OUT_NCHW, OUT_CNHW, OUT_NHWC, OUT_PLANES = 0, 1, 2, 3
CT_NEWEST_FIRST, CT_OLDEST_FIRST = 0, 1
CT_STANDARD, CT_TRANSPOSED = 0, 1


class Crop(C.Structure):
    _fields_ = [("data", C.c_void_p), ("width", C.c_int32), ("height", C.c_int32),
                ("pitch", C.c_int32), ("reserved", C.c_int32)]


class Op(C.Structure):
    _fields_ = [("kind", C.c_int32), ("perm", C.c_int32 * 4), ("v", C.c_float * 4)]


class Pipeline(C.Structure):
    _fields_ = [("src_type", C.c_int32), ("dst_width", C.c_int32), ("dst_height", C.c_int32),
                ("aspect_mode", C.c_int32), ("interp_mode", C.c_int32), ("fp_contract", C.c_int32),
                ("background", C.c_float * 4), ("n_ops", C.c_int32), ("ops", Op * MAX_OPS),
                ("out_layout", C.c_int32), ("dst_type", C.c_int32), ("out", C.c_void_p),
                ("out_plane_stride", C.c_int64), ("out_row_pitch", C.c_int64), ("yuv_standard", C.c_int32),
                ("u8_cast", C.c_int32)]


class Parent(C.Structure):
    _fields_ = [("datastart", C.c_void_p), ("whole_width", C.c_int32), ("whole_height", C.c_int32)]


class Plane(C.Structure):
    _fields_ = [("data", C.c_void_p), ("pitch_bytes", C.c_int64)]


class Warp(C.Structure):
    _fields_ = [("type", C.c_int32), ("m", C.c_float * 9)]


class Rect(C.Structure):
    _fields_ = [("x", C.c_int32), ("y", C.c_int32), ("width", C.c_int32), ("height", C.c_int32)]


# every symbol include/cvgs_b200.h declares: name -> (restype, argtypes)
SYMBOLS = {
    "cvgs_b200_version": (C.c_int, []),
    "cvgs_b200_last_error": (C.c_char_p, []),
    "cvgs_b200_preproc_launch": (C.c_int, [C.POINTER(Crop), C.c_int32, C.c_int32, C.POINTER(Pipeline), C.c_void_p]),
    "cvgs_b200_preproc_launch_ex": (C.c_int, [C.POINTER(Crop), C.POINTER(Parent), C.c_int32, C.c_int32,
                                              C.POINTER(Pipeline), C.c_void_p]),
    "cvgs_b200_preproc_launch_sequence_ex": (C.c_int, [C.POINTER(C.POINTER(Crop)), C.POINTER(C.POINTER(Parent)),
                                                       C.POINTER(C.c_int32), C.POINTER(C.c_int32),
                                                       C.POINTER(C.POINTER(Pipeline)), C.c_int32, C.c_int32,
                                                       C.c_void_p]),
    "cvgs_b200_warp_launch": (C.c_int, [C.POINTER(Crop), C.POINTER(Warp), C.c_int32, C.c_int32, C.POINTER(Pipeline),
                                        C.c_void_p]),
    "cvgs_b200_preproc_host": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.POINTER(Rect), C.c_int32,
                                         C.c_int32, C.POINTER(Pipeline), C.c_void_p, C.c_void_p]),
    "cvgs_b200_preproc_launch_sequence": (C.c_int, [C.POINTER(C.POINTER(Crop)), C.POINTER(C.c_int32),
                                                    C.POINTER(C.c_int32), C.POINTER(C.POINTER(Pipeline)), C.c_int32,
                                                    C.c_int32, C.c_void_p]),
    "cvgs_b200_preproc_host_sequence": (C.c_int, [C.POINTER(C.c_void_p), C.c_int32, C.c_int32, C.c_int32,
                                                  C.POINTER(C.POINTER(Rect)), C.POINTER(C.c_int32),
                                                  C.POINTER(C.c_int32), C.POINTER(C.POINTER(Pipeline)),
                                                  C.POINTER(C.c_void_p), C.c_int32, C.c_int32, C.c_void_p]),
    "cvgs_b200_set_kernel_variant": (C.c_int, [C.c_int]),
    "cvgs_b200_set_overlap": (C.c_int, [C.c_int]),
    "cvgs_b200_set_coalesce": (C.c_int, [C.c_int]),
    "cvgs_b200_set_host_upload": (C.c_int, [C.c_int]),
    "cvgs_b200_preproc_launch_rects": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.POINTER(Rect), C.c_int32,
                                                 C.c_int32, C.POINTER(Pipeline), C.c_void_p]),
    "cvgs_b200_preproc_launch_replicated": (C.c_int, [C.POINTER(Crop), C.POINTER(Parent), C.c_int32, C.c_int32,
                                                      C.POINTER(Pipeline), C.POINTER(C.c_void_p), C.c_int32, C.c_void_p]),
    "cvgs_b200_dev_alloc": (C.c_int, [C.POINTER(C.c_void_p), C.c_uint64]),
    "cvgs_b200_dev_free": (C.c_int, [C.c_void_p]),
    "cvgs_b200_ipc_export": (C.c_int, [C.c_void_p, C.c_void_p]),
    "cvgs_b200_ipc_open": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p)]),
    "cvgs_b200_ipc_close": (C.c_int, [C.c_void_p]),
    "cvgs_b200_launch_count": (C.c_int64, []),
    "cvgs_b200_debug_host_profile": (C.c_int, [C.POINTER(C.c_double), C.c_int]),
    "cvgs_b200_debug_fast_div": (C.c_uint32, [C.c_uint32, C.c_uint32]),
    "cvgs_b200_debug_warp_mode": (C.c_int, [C.POINTER(C.c_float), C.c_int32, C.c_int32, C.c_int32]),
    "cvgs_b200_debug_chain_kind": (C.c_int, [C.c_void_p, C.POINTER(C.c_float)]),
    "cvgs_b200_debug_host_bytes": (C.c_int, [C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.c_int]),
    "cvgs_b200_debug_program": (C.c_int, [C.POINTER(Pipeline), C.POINTER(C.c_float)]),
    "cvgs_b200_debug_plan": (C.c_int, [C.POINTER(Crop), C.c_int32, C.c_int32, C.POINTER(Pipeline), C.c_int32, C.c_int32,
                                       C.c_int32, C.POINTER(C.c_int64)]),
    "cvgs_b200_debug_overlap_query": (C.c_int, [C.c_void_p, C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint64]),
    "cvgs_b200_debug_division_sweep": (C.c_int, [C.c_float, C.POINTER(C.c_ulonglong), C.POINTER(C.c_uint),
                                                 C.POINTER(C.c_float)]),
    "cvgs_b200_ct_create": (C.c_int, [C.POINTER(C.c_void_p), C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                      C.c_int32, C.c_int32]),
    "cvgs_b200_ct_create_ex": (C.c_int, [C.POINTER(C.c_void_p)] + [C.c_int32] * 8),
    "cvgs_b200_ct_update": (C.c_int, [C.c_void_p, C.POINTER(Crop), C.POINTER(Pipeline), C.c_void_p]),
    "cvgs_b200_ct_data": (C.c_void_p, [C.c_void_p]),
    "cvgs_b200_ct_destroy": (C.c_int, [C.c_void_p]),
}

_lib = None


def load() -> C.CDLL:
    """Load libcvgs_b200.so (once) and bind every declared symbol."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: the CUDA library is the only implementation of this path "
            "(no CPU fallback). Build it with __graft_entry__.build().")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)  # AttributeError here = header and library disagree
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


class CvgsError(RuntimeError):
    """Mirrors the std::runtime_error the reference throws from gpuErrchk (fkl utils.h:42-60)."""


def check(rc: int) -> None:
    if rc != 0:
        msg = load().cvgs_b200_last_error().decode("utf-8", "replace")
        raise CvgsError(f"cvgs_b200 error {rc}: {msg}")
