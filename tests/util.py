"""Shared test helpers: oracle loader, synthetic workloads, comparisons.

The oracle (oracle/liboracle.so) is loaded ONLY here, in bench.py's cpu_baseline / reference legs and in
__graft_entry__.smoke(); the product never touches it.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass
from typing import List, Sequence, Tuple

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

from cvgpuspeedup_b200 import _abi, marshal  # noqa: E402  (struct layouts are those of include/cvgs_b200.h)

_ORACLE = None


def oracle_lib() -> C.CDLL:
    global _ORACLE
    if _ORACLE is None:
        path = os.path.join(ROOT, "oracle", "liboracle.so")
        if not os.path.exists(path):
            subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "liboracle.so"])
        lib = C.CDLL(path)
        lib.oracle_preproc.restype = C.c_int
        lib.oracle_preproc.argtypes = [C.POINTER(_abi.Crop), C.c_int, C.c_int, C.POINTER(_abi.Pipeline), C.c_int]

        class Geom(C.Structure):
            _fields_ = [("fx", C.c_float), ("fy", C.c_float), ("x1", C.c_int), ("y1", C.c_int), ("x2", C.c_int),
                        ("y2", C.c_int)]

        lib.Geom = Geom
        lib.oracle_resize_geometry.restype = None
        lib.oracle_resize_geometry.argtypes = [C.c_int] * 5 + [C.POINTER(Geom)]
        lib.oracle_ct_create.restype = C.c_void_p
        lib.oracle_ct_create.argtypes = [C.c_int] * 6
        lib.oracle_ct_create_ex.restype = C.c_void_p
        lib.oracle_ct_create_ex.argtypes = [C.c_int] * 7
        lib.oracle_ct_destroy.argtypes = [C.c_void_p]
        lib.oracle_ct_data.restype = C.POINTER(C.c_float)
        lib.oracle_ct_data.argtypes = [C.c_void_p]
        lib.oracle_ct_temp.restype = C.POINTER(C.c_float)
        lib.oracle_ct_temp.argtypes = [C.c_void_p]
        lib.oracle_ct_update.restype = C.c_int
        lib.oracle_ct_update.argtypes = [C.c_void_p, C.POINTER(_abi.Crop), C.POINTER(_abi.Pipeline), C.c_int]
        lib.oracle_warp.restype = C.c_int
        lib.oracle_warp.argtypes = [C.POINTER(_abi.Crop), C.POINTER(_abi.Warp), C.c_int, C.c_int,
                                    C.POINTER(_abi.Pipeline), C.c_int]
        lib.oracle_has_fma.restype = C.c_int
        lib.oracle_max_threads.restype = C.c_int
        assert lib.oracle_has_fma() == 1, "host CPU lacks FMA: the oracle would not be exact"
        _ORACLE = lib
    return _ORACLE


Rect = Tuple[int, int, int, int]  # x, y, w, h


@dataclass
class Workload:
    """A synthetic batch: one source image (HxWx3 uint8, row pitch may exceed 3*W), crops and a chain."""
    name: str
    image: np.ndarray          # [H, pitch] uint8 backing store (pitch >= 3*W)
    width: int
    height: int
    rects: List[Rect]
    dsize: Tuple[int, int]     # (W, H) like cv::Size
    ops: list                  # list of (kind, (v0,v1,v2)) / ("reorder", (2,1,0))
    aspect: int = _abi.IGNORE_AR
    background: Tuple[float, float, float] = (0.0, 0.0, 0.0)

    @property
    def pitch(self) -> int:
        return self.image.shape[1]


def make_image(rng, width, height, pitch=None, smooth=False) -> np.ndarray:
    pitch = pitch or 3 * width
    img = np.zeros((height, pitch), dtype=np.uint8)
    if smooth:
        yy, xx = np.mgrid[0:height, 0:width]
        base = (127.5 + 127.5 * np.sin(xx / 17.0) * np.cos(yy / 23.0))
        px = np.stack([base, 255 - base, (base * 0.5 + 60)], axis=-1)
        px = np.clip(px + rng.normal(0, 2, px.shape), 0, 255).astype(np.uint8)
    else:
        px = rng.integers(0, 256, size=(height, width, 3), dtype=np.uint8)
    img[:, :3 * width] = px.reshape(height, 3 * width)
    if pitch > 3 * width:  # padding bytes must never influence results: fill with noise
        img[:, 3 * width:] = rng.integers(0, 256, size=(height, pitch - 3 * width), dtype=np.uint8)
    return img


OPS_C2 = [("reorder", (2, 1, 0)), ("mul", (0.3, 0.3, 0.3)), ("sub", (1.0, 4.0, 3.2)), ("div", (3.2, 0.6, 11.8))]
OPS_C1 = [("mul", (0.5, 0.5, 0.5)), ("sub", (1.0, 4.0, 6.0)), ("div", (2.0, 8.0, 1.0))]
_MEAN, _STD = (0.485, 0.456, 0.406), (0.229, 0.224, 0.225)
OPS_C3 = [("reorder", (2, 1, 0)), ("mul", (1 / 255.0,) * 3), ("sub", _MEAN), ("div", _STD)]


def workload_c1(seed=1, smooth=False) -> Workload:
    """BASELINE config 1: one 640x480 CV_8UC3 crop -> 64x128, alpha 0.5, sub/div (README.md:71-72,88)."""
    rng = np.random.default_rng(seed)
    return Workload("c1", make_image(rng, 640, 480, smooth=smooth), 640, 480, [(0, 0, 640, 480)], (64, 128), OPS_C1)


def workload_c2(seed=2, n=50, pitch=6144, ref_shape=False, frame=(1920, 1080)) -> Workload:
    """BASELINE config 2: n crops of mixed size from a 1080p frame -> 64x128 + normalise + split."""
    rng = np.random.default_rng(seed)
    fw, fh = frame
    img = make_image(rng, fw, fh, pitch)
    rects = []
    for i in range(n):
        if ref_shape:  # tests/batchresize/test_batchresize_x_split3D.cu:69-78
            rects.append((i, i, 60, 120))
        else:
            w = int(rng.integers(24, 257))
            h = min(2 * w, fh)
            rects.append((int(rng.integers(0, fw - w + 1)), int(rng.integers(0, fh - h + 1)), w, h))
    return Workload("c2", img, fw, fh, rects, (64, 128), OPS_C2)


def workload_c3(seed=3, n=256, frame=(3840, 2160), dsize=(224, 224), lo=224, hi=896) -> Workload:
    """BASELINE config 3: n crops from a 4K frame -> 224x224, BGR2RGB, ImageNet mean/std, NCHW."""
    rng = np.random.default_rng(seed)
    fw, fh = frame
    img = make_image(rng, fw, fh)
    rects = []
    for _ in range(n):
        w, h = int(rng.integers(lo, hi + 1)), int(rng.integers(lo, hi + 1))
        rects.append((int(rng.integers(0, fw - w + 1)), int(rng.integers(0, fh - h + 1)), w, h))
    return Workload("c3", img, fw, fh, rects, dsize, OPS_C3)


from cvgpuspeedup_b200.marshal import OP_KINDS as _KIND, make_pipeline  # noqa: E402,F401


def out_channels(src_type, ops) -> int:
    """Channels of the pixel a chain ends with (the alpha / gray conversions change the count)."""
    nc = channels_of(src_type)
    for k, _ in ops:
        nc = {"add_alpha": 4, "drop_alpha": 3, "gray": 1}.get(k, nc)
    return nc


def px_bytes_of(src_type) -> int:
    return {_abi.CVGS_8UC3: 3, _abi.CVGS_8UC4: 4, _abi.CVGS_16UC3: 6, _abi.CVGS_16SC3: 6, _abi.CVGS_NV12: 1}.get(src_type, 8)


def channels_of(src_type) -> int:
    return 4 if src_type in (_abi.CVGS_8UC4, _abi.CVGS_16UC4, _abi.CVGS_16SC4) else 3


def out_shape(n_planes, dsize, layout, plane_stride=0, nc=3):
    W, H = dsize
    if plane_stride:
        if layout == _abi.OUT_CNHW:
            return (nc, n_planes, plane_stride)
        return (n_planes, plane_stride)
    if layout == _abi.OUT_NCHW:
        return (n_planes, nc, H, W)
    if layout == _abi.OUT_CNHW:
        return (nc, n_planes, H, W)
    return (n_planes, H, W, nc)


def host_crops(image: np.ndarray, rects: Sequence[Rect], base_ptr: int | None = None, px_bytes: int = 3):
    """Crop descriptors pointing into `image` (host) or into a device copy at base_ptr with the same pitch.
    px_bytes = 3 for CV_8UC3, 6 for the 16-bit sources."""
    return marshal.crop_array(image.ctypes.data if base_ptr is None else base_ptr, image.shape[1], rects, px_bytes)


def host_parents(image: np.ndarray, width: int, height: int, n: int, base_ptr: int | None = None):
    """cvgs_parent_t per crop: every crop was cut from the one image (GpuMat::datastart + locateROI)."""
    return marshal.parent_array(image.ctypes.data if base_ptr is None else base_ptr, width, height, n)


def run_oracle(image, rects, dsize, ops, n_planes=None, used=None, nthreads=0, fill=np.nan, **pipe_kw) -> np.ndarray:
    """CPU oracle on host memory; returns the output tensor (shape by layout)."""
    lib = oracle_lib()
    n_planes = len(rects) if n_planes is None else n_planes
    used = len(rects) if used is None else used
    layout = pipe_kw.get("layout", _abi.OUT_NCHW)
    st = pipe_kw.get("src_type", _abi.CVGS_8UC3)
    out = np.full(out_shape(n_planes, dsize, layout, pipe_kw.get("plane_stride", 0), out_channels(st, ops)), fill, dtype=np.float32)
    p = make_pipeline(dsize, ops, out_ptr=out.ctypes.data, **pipe_kw)
    crops = host_crops(image, rects[:used], px_bytes=px_bytes_of(st))
    rc = lib.oracle_preproc(crops, n_planes, used, C.byref(p), nthreads)
    assert rc == 0, "oracle rejected the arguments"
    return out


def assert_bit_equal(a: np.ndarray, b: np.ndarray, what=""):
    """Bit-exact float comparison (NaN payloads included), with a useful message."""
    a32, b32 = np.ascontiguousarray(a).view(np.uint32), np.ascontiguousarray(b).view(np.uint32)
    if a32.shape != b32.shape:
        raise AssertionError(f"{what}: shape {a.shape} vs {b.shape}")
    bad = np.nonzero(a32 != b32)
    if bad[0].size:
        idx = tuple(int(d[0]) for d in bad)
        raise AssertionError(f"{what}: {bad[0].size} of {a32.size} values differ; first at {idx}: "
                             f"{a[idx]!r} vs {b[idx]!r}")


def ulp_diff(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    """Distance in units in the last place between float32 arrays (finite values)."""
    def key(x):
        i = np.ascontiguousarray(x, dtype=np.float32).view(np.int32).astype(np.int64)
        return np.where(i < 0, -(i & 0x7FFFFFFF), i)
    return np.abs(key(a) - key(b))
