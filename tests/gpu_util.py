"""GPU-side test helpers: run the product through its C-ABI, and run the reference's real kernel
(oracle/_ref/libfkref_N.so, built from /root/reference by oracle/Makefile) for bit-exact comparison."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np
import torch

from cvgpuspeedup_b200 import _abi
from tests import util

ROOT = util.ROOT


def device_image(image: np.ndarray) -> torch.Tensor:
    return torch.from_numpy(image).cuda()


def run_cvgs(image, rects, dsize, ops, n_planes=None, used=None, variant=0, fill=float("nan"), d_image=None,
             parents=None, **pipe_kw) -> np.ndarray:
    """One cvgs_b200_preproc_launch (or _launch_ex when parents = (width, height) of the image the crops were cut
    from) on cuda:0; returns the output tensor as numpy."""
    lib = _abi.load()
    n_planes = len(rects) if n_planes is None else n_planes
    used = len(rects) if used is None else used
    layout = pipe_kw.get("layout", _abi.OUT_NCHW)
    d_img = device_image(image) if d_image is None else d_image
    st = pipe_kw.get("src_type", _abi.CVGS_8UC3)
    shape = util.out_shape(n_planes, dsize, layout, pipe_kw.get("plane_stride", 0), util.out_channels(st, ops))
    d_out = torch.full(shape, fill, dtype=torch.float32, device="cuda")
    p = util.make_pipeline(dsize, ops, out_ptr=d_out.data_ptr(), **pipe_kw)
    crops = util.host_crops(image, rects[:used], base_ptr=d_img.data_ptr(), px_bytes=util.px_bytes_of(st))
    prev = lib.cvgs_b200_set_kernel_variant(variant)
    try:
        if parents is None:
            _abi.check(lib.cvgs_b200_preproc_launch(crops, n_planes, used, C.byref(p),
                                                    torch.cuda.current_stream().cuda_stream))
        else:
            par = util.host_parents(image, parents[0], parents[1], used, base_ptr=d_img.data_ptr())
            _abi.check(lib.cvgs_b200_preproc_launch_ex(crops, par, n_planes, used, C.byref(p),
                                                       torch.cuda.current_stream().cuda_stream))
    finally:
        lib.cvgs_b200_set_kernel_variant(prev)
    torch.cuda.synchronize()
    return d_out.cpu().numpy()


_FKREF = {}


def fkref_lib(batch: int):
    if batch not in _FKREF:
        path = os.path.join(ROOT, "oracle", "_ref", f"libfkref_{batch}.so")
        if not os.path.exists(path):
            return None
        lib = C.CDLL(path)
        fn = getattr(lib, f"fkref_preproc_{batch}")
        fn.restype = C.c_int
        fn.argtypes = [C.POINTER(C.c_void_p), C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int), C.c_int,
                       C.c_int, C.c_int, C.c_int, C.POINTER(C.c_float), C.c_int, C.POINTER(C.c_float),
                       C.POINTER(C.c_float), C.POINTER(C.c_float), C.c_void_p, C.c_void_p]
        err = getattr(lib, f"fkref_last_error_{batch}")
        err.restype = C.c_char_p
        _FKREF[batch] = (fn, err)
    return _FKREF[batch]


def fkref_args(d_img_ptr, pitch, rects, used):
    n = max(1, used)
    ptrs = (C.c_void_p * n)(*[d_img_ptr + y * pitch + 3 * x for (x, y, w, h) in rects[:used]])
    ws = (C.c_int * n)(*[r[2] for r in rects[:used]])
    hs = (C.c_int * n)(*[r[3] for r in rects[:used]])
    ps = (C.c_int * n)(*[pitch] * used)
    return ptrs, ws, hs, ps


def run_fkref(image, rects, dsize, swap, mul, sub, div, aspect=_abi.IGNORE_AR, bg=(0, 0, 0), batch=16, used=None,
              d_image=None, src_type=_abi.CVGS_8UC3) -> np.ndarray:
    """The reference's own fused kernel: resize -> [RGB2BGR] -> Mul -> Sub -> Div -> TensorSplit, BATCH = batch
    (a template parameter there).  Returns [batch, 3, H, W]."""
    lib = fkref_lib(batch)
    assert lib is not None, "oracle/_ref not built"
    fn, err = lib
    used = len(rects) if used is None else used
    assert used <= batch
    d_img = device_image(image) if d_image is None else d_image
    d_out = torch.full((batch, util.channels_of(src_type), dsize[1], dsize[0]), float("nan"), dtype=torch.float32,
                       device="cuda")
    f3 = lambda v: (C.c_float * 4)(*(tuple(v) + (0.0,) * (4 - len(v))))  # noqa: E731
    if src_type != _abi.CVGS_8UC3:  # 16-bit and 4-channel instantiations exist in the BATCH=16 library only
        fn16 = getattr(C.CDLL(os.path.join(ROOT, "oracle", "_ref", f"libfkref_{batch}.so")), f"fkref_preproc16_{batch}")
        fn16.restype = C.c_int
        fn16.argtypes = [C.c_int] + fn.argtypes
        pitch = image.shape[1]
        n = max(1, used)
        pb = util.px_bytes_of(src_type)
        ptrs = (C.c_void_p * n)(*[d_img.data_ptr() + y * pitch + pb * x for (x, y, w, h) in rects[:used]])
        _, ws, hs, ps = fkref_args(d_img.data_ptr(), pitch, rects, used)
        rc = fn16(src_type, ptrs, ws, hs, ps, used, dsize[0], dsize[1], aspect, f3(bg), int(swap), f3(mul), f3(sub), f3(div),
                  d_out.data_ptr(), torch.cuda.current_stream().cuda_stream)
        assert rc == 0, err().decode()
        torch.cuda.synchronize()
        return d_out.cpu().numpy()
    ptrs, ws, hs, ps = fkref_args(d_img.data_ptr(), image.shape[1], rects, used)
    rc = fn(ptrs, ws, hs, ps, used, dsize[0], dsize[1], aspect, f3(bg), int(swap), f3(mul), f3(sub), f3(div),
            d_out.data_ptr(), torch.cuda.current_stream().cuda_stream)
    assert rc == 0, err().decode()
    torch.cuda.synchronize()
    return d_out.cpu().numpy()


_CHAIN = None


def chain_lib():
    """oracle/libchain.so: the restated multi-kernel 'OpenCV-CUDA-equivalent' chain (baseline M)."""
    global _CHAIN
    if _CHAIN is None:
        path = os.path.join(ROOT, "oracle", "libchain.so")
        if not os.path.exists(path):
            return None
        lib = C.CDLL(path)
        P, I = C.POINTER, C.c_int
        lib.chain_workspace_bytes.restype = C.c_size_t
        lib.chain_workspace_bytes.argtypes = [I, I]
        lib.chain_preproc.restype = I
        lib.chain_preproc.argtypes = [P(C.c_void_p), P(I), P(I), P(I), I, I, I, I, P(C.c_float), P(C.c_float), P(C.c_float),
                                      C.c_void_p, C.c_void_p, C.c_void_p]
        lib.chain_preproc_sequence.restype = I
        lib.chain_preproc_sequence.argtypes = [P(P(C.c_void_p)), P(P(I)), P(P(I)), P(P(I)), I, I, I, I, P(C.c_float),
                                               P(C.c_float), P(C.c_float), P(C.c_void_p), C.c_void_p, I, I, C.c_void_p]
        _CHAIN = lib
    return _CHAIN


def run_chain(image, rects, dsize, swap, mul, sub, div, d_image=None):
    """resize(8U) -> convertTo(alpha) -> [swap] -> subtract -> divide -> split, one launch per step and crop.
    Returns ([n, 3, H, W] numpy, number of kernel launches)."""
    lib = chain_lib()
    assert lib is not None, "oracle/libchain.so not built"
    d_img = device_image(image) if d_image is None else d_image
    n = len(rects)
    out = torch.full((n, 3, dsize[1], dsize[0]), float("nan"), dtype=torch.float32, device="cuda")
    ws = torch.empty(int(lib.chain_workspace_bytes(dsize[0], dsize[1])) + 512, dtype=torch.uint8, device="cuda")
    ptrs, w_, h_, p_ = fkref_args(d_img.data_ptr(), image.shape[1], rects, n)
    f3 = lambda v: (C.c_float * 3)(*v)  # noqa: E731
    launches = lib.chain_preproc(ptrs, w_, h_, p_, n, dsize[0], dsize[1], int(swap), f3(mul), f3(sub), f3(div),
                                 out.data_ptr(), ws.data_ptr(), torch.cuda.current_stream().cuda_stream)
    assert launches > 0
    torch.cuda.synchronize()
    return out.cpu().numpy(), launches


def run_fkref_nv12(frame, width, height, dsize, standard, mul, sub, div, d_frame=None):
    """The reference's ReadYUV<NV12> + ConvertYUVToRGB + Resize + Mul/Sub/Div + TensorSplit on one frame
    (oracle/_ref/libfkref_16.so, built with -DFKREF_NV12).  frame: [height + (height + 1) // 2, pitch] uint8."""
    path = os.path.join(ROOT, "oracle", "_ref", "libfkref_16.so")
    lib = C.CDLL(path)
    fn = lib.fkref_nv12_16
    fn.restype = C.c_int
    fn.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_float),
                   C.POINTER(C.c_float), C.POINTER(C.c_float), C.c_void_p, C.c_void_p]
    d = device_image(frame) if d_frame is None else d_frame
    out = torch.full((3, dsize[1], dsize[0]), float("nan"), dtype=torch.float32, device="cuda")
    f3 = lambda v: (C.c_float * 3)(*v)  # noqa: E731
    rc = fn(standard, d.data_ptr(), width, height, frame.shape[1], dsize[0], dsize[1], f3(mul), f3(sub), f3(div),
            out.data_ptr(), torch.cuda.current_stream().cuda_stream)
    assert rc == 0
    torch.cuda.synchronize()
    return out.cpu().numpy()


def run_fkref_warp(image, width, height, warp_type, inverse, dsize, mul=None, d_image=None):
    """The reference's fk::Warping<WT, PerThreadRead<_2D, uchar3>> on one image (oracle/_ref/libfkref_16.so, built with
    -DFKREF_WARP): mul given -> Mul + TensorSplit, float [3, H, W]; mul None -> fk::Cast<float3, uchar3> + packed write,
    uint8 [H, W, 3].  `inverse`: nine floats, destination -> source."""
    path = os.path.join(ROOT, "oracle", "_ref", "libfkref_16.so")
    lib = C.CDLL(path)
    fn = lib.fkref_warp_16
    fn.restype = C.c_int
    fn.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_float), C.c_int, C.c_int,
                   C.POINTER(C.c_float), C.c_void_p, C.c_int, C.c_void_p]
    d = device_image(image) if d_image is None else d_image
    m = (C.c_float * 9)(*[float(v) for v in inverse])
    if mul is not None:
        out = torch.full((3, dsize[1], dsize[0]), float("nan"), dtype=torch.float32, device="cuda")
        rc = fn(warp_type, 0, d.data_ptr(), width, height, image.shape[1], m, dsize[0], dsize[1], (C.c_float * 3)(*mul),
                out.data_ptr(), 0, torch.cuda.current_stream().cuda_stream)
    else:
        out = torch.full((dsize[1], dsize[0], 3), 77, dtype=torch.uint8, device="cuda")
        rc = fn(warp_type, 1, d.data_ptr(), width, height, image.shape[1], m, dsize[0], dsize[1], None,
                out.data_ptr(), 3 * dsize[0], torch.cuda.current_stream().cuda_stream)
    assert rc == 0
    torch.cuda.synchronize()
    return out.cpu().numpy()


CVT_CODES = {  # cv:: / fk:: ColorConversionCodes -> (source channels, ops of the C-ABI)
    0: (3, [("add_alpha", (255.0,))]),                                  # BGR2BGRA / RGB2RGBA
    1: (4, [("drop_alpha", ())]),                                       # BGRA2BGR / RGBA2RGB
    2: (3, [("reorder", (2, 1, 0)), ("add_alpha", (255.0,))]),          # BGR2RGBA / RGB2BGRA
    3: (4, [("reorder", (2, 1, 0, 3)), ("drop_alpha", ())]),            # RGBA2BGR / BGRA2RGB
    6: (3, [("reorder", (2, 1, 0)), ("gray", ())]),                     # BGR2GRAY
    7: (3, [("gray", (1,))]),                                           # RGB2GRAY (gray (1,): y * 0.587 first)
    10: (4, [("reorder", (2, 1, 0, 3)), ("gray", ())]),                 # BGRA2GRAY
    11: (4, [("gray", (1,))]),                                          # RGBA2GRAY
}


def run_fkref_cvt(code, image, width, height, dsize, mul, sub, d_image=None):
    """The reference's Resize + ColorConversion<code> + Mul + Sub + write on one image (oracle/_ref/libfkref_16.so,
    -DFKREF_CVT).  Returns [C, H, W] float32 (C = channels after the conversion)."""
    path = os.path.join(ROOT, "oracle", "_ref", "libfkref_16.so")
    lib = C.CDLL(path)
    fn = lib.fkref_cvt_16
    fn.restype = C.c_int
    fn.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_float),
                   C.POINTER(C.c_float), C.c_void_p, C.c_void_p]
    nc_out = {0: 4, 2: 4, 1: 3, 3: 3}.get(code, 1)
    d = device_image(image) if d_image is None else d_image
    out = torch.full((nc_out, dsize[1], dsize[0]), float("nan"), dtype=torch.float32, device="cuda")
    f4 = lambda v: (C.c_float * 4)(*(tuple(v) + (0.0,) * 4)[:4])  # noqa: E731
    rc = fn(code, d.data_ptr(), width, height, image.shape[1], dsize[0], dsize[1], f4(mul), f4(sub), out.data_ptr(),
            torch.cuda.current_stream().cuda_stream)
    assert rc == 0
    torch.cuda.synchronize()
    return out.cpu().numpy()


def run_fkref_yuv(fmt, frame, width, height, dsize, standard, mul, sub, div, d_frame=None):
    """The reference's ReadYUV<fmt> + ConvertYUVToRGB + Resize + Mul/Sub/Div + TensorSplit on one frame
    (oracle/_ref/libfkref_16.so, -DFKREF_YUV; standards 0 and 3).  fmt: _abi.CVGS_NV21 / P010 / P210 / Y210."""
    path = os.path.join(ROOT, "oracle", "_ref", "libfkref_16.so")
    lib = C.CDLL(path)
    fn = lib.fkref_yuv_16
    fn.restype = C.c_int
    fn.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_float),
                   C.POINTER(C.c_float), C.POINTER(C.c_float), C.c_void_p, C.c_void_p]
    d = device_image(frame) if d_frame is None else d_frame
    out = torch.full((3, dsize[1], dsize[0]), float("nan"), dtype=torch.float32, device="cuda")
    f3 = lambda v: (C.c_float * 3)(*v)  # noqa: E731
    rc = fn(fmt & 0xF, standard, d.data_ptr(), width, height, frame.shape[1], dsize[0], dsize[1], f3(mul), f3(sub), f3(div),
            out.data_ptr(), torch.cuda.current_stream().cuda_stream)
    assert rc == 0
    torch.cuda.synchronize()
    return out.cpu().numpy()


def run_fkref_warp_typed(image, width, height, src_type, warp_type, inverse, dsize, mul, d_image=None):
    """fk::Warping on a CV_8UC4 / CV_16UC3 / CV_16SC4 image -> Mul -> TensorSplit (oracle/_ref/libfkref_16.so)."""
    path = os.path.join(ROOT, "oracle", "_ref", "libfkref_16.so")
    lib = C.CDLL(path)
    fn = lib.fkref_warp_typed_16
    fn.restype = C.c_int
    fn.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_float), C.c_int, C.c_int,
                   C.POINTER(C.c_float), C.c_void_p, C.c_void_p]
    d = device_image(image) if d_image is None else d_image
    nc = util.channels_of(src_type)
    out = torch.full((nc, dsize[1], dsize[0]), float("nan"), dtype=torch.float32, device="cuda")
    m = (C.c_float * 9)(*[float(v) for v in inverse])
    rc = fn(warp_type, src_type, d.data_ptr(), width, height, image.shape[1], m, dsize[0], dsize[1],
            (C.c_float * 4)(*(tuple(mul) + (0.0,) * 4)[:4]), out.data_ptr(), torch.cuda.current_stream().cuda_stream)
    assert rc == 0
    torch.cuda.synchronize()
    return out.cpu().numpy()
