"""Host-side mirror of the reference's operator interface for the hot path.

Same names and argument meaning as ``namespace cvGS`` (reference include/cvGPUSpeedup.cuh):
``resize``, ``cvtColor``, ``multiply``, ``subtract``, ``divide``, ``add``, ``convertTo``,
``split``, ``splitT``, ``write``, ``executeOperations`` and ``CircularTensor``.  The functions
build small descriptor objects (the reference builds operation-struct PODs) and
``executeOperations`` turns the chain into ONE call of the C-ABI, i.e. one kernel launch.

torch is used only as the owner of device memory and streams (``GpuMat`` wraps a uint8 CUDA
tensor the way cv::cuda::GpuMat wraps a device allocation).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import List, Optional, Sequence, Tuple

from . import _abi
from ._abi import (CT_NEWEST_FIRST, CT_OLDEST_FIRST, CT_STANDARD, CT_TRANSPOSED, FP_REFERENCE_FUSED,
                   FP_SEPARATE, IGNORE_AR, INTERP_FLOAT, INTERP_ROUND_U8, OUT_CNHW, OUT_NCHW, OUT_NHWC,
                   OUT_PLANES, PRESERVE_AR, PRESERVE_AR_LEFT, PRESERVE_AR_RN_EVEN, CvgsError)

# cv::ColorConversionCodes values used by the reference tests (cv2cuda_types.cuh:77-86)
COLOR_BGR2RGB = 4
COLOR_RGB2BGR = 4
INTER_LINEAR = 1
CV_8UC3 = _abi.CVGS_8UC3
CV_16UC3 = _abi.CVGS_16UC3
CV_16SC3 = _abi.CVGS_16SC3
CV_8UC4, CV_16UC4, CV_16SC4, CV_32FC4 = _abi.CVGS_8UC4, _abi.CVGS_16UC4, _abi.CVGS_16SC4, _abi.CVGS_32FC4
COLOR_RGBA2BGRA = COLOR_BGRA2RGBA = 5
COLOR_BGR2BGRA = COLOR_RGB2RGBA = 0
COLOR_BGRA2BGR = COLOR_RGBA2RGB = 1
COLOR_BGR2RGBA = COLOR_RGB2BGRA = 2
COLOR_RGBA2BGR = COLOR_BGRA2RGB = 3
COLOR_BGR2GRAY, COLOR_RGB2GRAY, COLOR_BGRA2GRAY, COLOR_RGBA2GRAY = 6, 7, 10, 11
CV_32FC3 = _abi.CVGS_32FC3


class GpuMat:
    """Minimal cv::cuda::GpuMat stand-in: the wrapper only ever touches data/cols/rows/step
    (reference include/cvGPUSpeedup.cuh:36,42,69)."""

    __slots__ = ("data", "cols", "rows", "step", "_owner", "datastart", "whole", "elem_size")

    def __init__(self, data: int, cols: int, rows: int, step: int, owner=None, datastart: Optional[int] = None,
                 whole: Optional[Tuple[int, int]] = None, elem_size: int = 3):
        self.data, self.cols, self.rows, self.step, self._owner = int(data), int(cols), int(rows), int(step), owner
        self.elem_size = int(elem_size)  # GpuMat::elemSize(): 3 for CV_8UC3, 4 for CV_8UC4, 6 / 8 for the 16-bit types
        # cv::cuda::GpuMat::datastart / locateROI(wholeSize, ofs): the image this header was cut from
        self.datastart = int(data) if datastart is None else int(datastart)
        self.whole = (int(cols), int(rows)) if whole is None else (int(whole[0]), int(whole[1]))

    @classmethod
    def from_tensor(cls, t) -> "GpuMat":
        """t: uint8 CUDA tensor [H, W, 3] whose rows are contiguous (stride(1)==3, stride(2)==1)."""
        if t.dim() != 3 or t.shape[2] != 3 or t.stride(2) != 1 or t.stride(1) != 3 or not t.is_cuda:
            raise ValueError("expected a CUDA uint8 HxWx3 tensor with packed pixels")
        return cls(t.data_ptr(), t.shape[1], t.shape[0], t.stride(0), owner=t)

    def roi(self, x: int, y: int, w: int, h: int) -> "GpuMat":
        """d_input(cv::Rect(x, y, w, h))"""
        if x < 0 or y < 0 or w <= 0 or h <= 0 or x + w > self.cols or y + h > self.rows:
            raise ValueError("ROI outside the image")
        return GpuMat(self.data + y * self.step + self.elem_size * x, w, h, self.step, owner=self._owner,
                      datastart=self.datastart, whole=self.whole, elem_size=self.elem_size)


def crop(image: "GpuMat", rects) -> List["GpuMat"]:
    """cvGS::crop(readOfImage, rect) / crop(readOfImage, rects) (reference include/cvGPUSpeedup.cuh:247-265,444 ->
    fk::Crop, crop.cuh:23-55): the rectangles (x, y, w, h) of one image as the crop list resize() / executeOperations take.
    Every crop remembers the image it was cut from, so the launch stages them through the image's cached tensor maps."""
    if rects and isinstance(rects[0], (int, float)):
        rects = [rects]
    return [image.roi(int(x), int(y), int(w), int(h)) for (x, y, w, h) in rects]


def _scalar3(s) -> Tuple[float, ...]:
    """cv::Scalar: up to four values (the fourth is used by 4-channel pipelines only)."""
    if isinstance(s, (int, float)):
        return (float(s),) * 4
    s = tuple(float(v) for v in s)
    if len(s) < 3:
        raise ValueError("a cv::Scalar with at least 3 values is required")
    return (s + (0.0,))[:4]


@dataclass
class _Resize:
    crops: List[GpuMat]
    dsize: Tuple[int, int]  # (width, height) like cv::Size
    used: int
    background: Tuple[float, float, float]
    aspect: int
    src_type: int = _abi.CVGS_8UC3
    yuv_standard: int = 0


@dataclass
class _Op:
    kind: int
    v: Tuple[float, ...] = (0.0, 0.0, 0.0, 0.0)
    perm: Tuple[int, ...] = (0, 1, 2, 3)


@dataclass
class _Write:
    out_ptr: int
    layout: int
    plane_stride: int = 0
    owner: object = None
    planes: object = None  # OUT_PLANES: ctypes array of _abi.Plane kept alive with the op
    dst_type: int = 0      # _abi.CVGS_8UC3: packed 8-bit image
    row_pitch: int = 0
    u8_cast: int = 0


@dataclass
class _Warp:
    images: List[GpuMat]
    inverse: list          # per image (type, nine float32) -- destination -> source
    dsize: Tuple[int, int]
    used: int
    background: Tuple[float, float, float]
    src_type: int = _abi.CVGS_8UC3


def resize(crops: Sequence[GpuMat], dsize: Tuple[int, int], usedPlanes: Optional[int] = None,
           backgroundValue=(0.0, 0.0, 0.0), aspect: int = IGNORE_AR, src_type: int = _abi.CVGS_8UC3) -> _Resize:
    """cvGS::resize<CV_8UC3, INTER_LINEAR, N, AR>(array<GpuMat,N>, Size, usedPlanes, bg)
    (reference include/cvGPUSpeedup.cuh:218-245)."""
    crops = list(crops)
    return _Resize(crops, (int(dsize[0]), int(dsize[1])), len(crops) if usedPlanes is None else int(usedPlanes),
                   _scalar3(backgroundValue), int(aspect), int(src_type))


def resize_nv12(frames: Sequence[GpuMat], dsize: Tuple[int, int], standard: int = _abi.YUV_BT709_FULL,
                usedPlanes: Optional[int] = None, backgroundValue=(0.0, 0.0, 0.0), aspect: int = IGNORE_AR,
                fmt: int = _abi.CVGS_NV12) -> _Resize:
    """fk::Resize<INTER_LINEAR>::build(fk::fuse(Read<ReadYUV<NV12>>, Unary<ConvertYUVToRGB<NV12, range, primaries, false,
    float3>>), dsize) for a batch of NV12 frames (reference color_conversion.cuh:235-362, tests/resize/
    test_fused_resize.cu:73-76).  Each GpuMat describes the luma plane (cols x rows, step); the interleaved UV plane
    follows it at data + step * rows.  The chain after it sees float RGB.  `fmt`: _abi.CVGS_NV12 / NV21 / P010 / P210 /
    Y210 (include/cvgs_b200.h)."""
    r = resize(frames, dsize, usedPlanes, backgroundValue, aspect, int(fmt))
    r.yuv_standard = int(standard)
    return r


WARP_AFFINE, WARP_PERSPECTIVE = _abi.WARP_AFFINE, _abi.WARP_PERSPECTIVE


def invert_warp_matrix(m, warp_type: int):
    """What cvGS::warp does to the user's matrix on the host (reference include/cvGPUSpeedup.cuh:266-283): affine 2x3
    through cv::invertAffineTransform, perspective 3x3 through Mat::inv(), both in double, then cast to float.
    Returns nine float32 (row-major; the third row of an affine matrix is unused)."""
    import numpy as np
    m = np.asarray(m, dtype=np.float64)
    out = np.zeros(9, dtype=np.float32)
    if warp_type == WARP_AFFINE:
        if m.shape != (2, 3):
            raise CvgsError("an affine warp takes a 2x3 matrix")
        # cv::invertAffineTransform (imgproc/src/imgwarp.cpp): closed form, D = 0 gives a zero matrix
        d = m[0, 0] * m[1, 1] - m[0, 1] * m[1, 0]
        d = 1.0 / d if d != 0 else 0.0
        a11, a22, a12, a21 = m[1, 1] * d, m[0, 0] * d, -m[0, 1] * d, -m[1, 0] * d
        b1 = -a11 * m[0, 2] - a12 * m[1, 2]
        b2 = -a21 * m[0, 2] - a22 * m[1, 2]
        out[:6] = np.array([a11, a12, b1, a21, a22, b2], dtype=np.float64).astype(np.float32)
    elif warp_type == WARP_PERSPECTIVE:
        if m.shape != (3, 3):
            raise CvgsError("a perspective warp takes a 3x3 matrix")
        out[:] = np.linalg.inv(m).astype(np.float32).ravel()
    else:
        raise CvgsError("warp type must be WARP_AFFINE or WARP_PERSPECTIVE")
    return out


def warp(images, matrices, dsize: Tuple[int, int], warp_type: int = WARP_AFFINE, usedPlanes: Optional[int] = None,
         backgroundValue=(0.0, 0.0, 0.0), src_type: int = _abi.CVGS_8UC3) -> _Warp:
    """cvGS::warp<WT, CV_8UC3[, N]>(GpuMat / array<GpuMat, N>, Mat / array<Mat, N>, Size[, usedPlanes, default])
    (reference include/cvGPUSpeedup.cuh:285-442): `matrices` are the forward transforms cv::warpAffine /
    cv::warpPerspective take; they are inverted here like the wrapper does."""
    if isinstance(images, GpuMat):
        images, matrices = [images], [matrices]
    images = list(images)
    inv = [invert_warp_matrix(m, warp_type) for m in matrices]
    if len(inv) != len(images):
        raise CvgsError("one matrix per image is required")
    return _Warp(images, [(int(warp_type), v) for v in inv], (int(dsize[0]), int(dsize[1])),
                 len(images) if usedPlanes is None else int(usedPlanes), _scalar3(backgroundValue), int(src_type))


def multiply(s) -> _Op:   # cvGS::multiply<CV_32FC3>(Scalar) :131
    return _Op(_abi.OP_MUL, _scalar3(s))


def subtract(s) -> _Op:   # cvGS::subtract :136
    return _Op(_abi.OP_SUB, _scalar3(s))


def divide(s) -> _Op:     # cvGS::divide :141
    return _Op(_abi.OP_DIV, _scalar3(s))


def add(s) -> _Op:        # cvGS::add :146
    return _Op(_abi.OP_ADD, _scalar3(s))


def convertTo(alpha: Optional[float] = None, beta: Optional[float] = None) -> List[_Op]:
    """cvGS::convertTo<CV_8UC3, CV_32FC3>([alpha[, beta]]) :74-129 -- the cast itself is implicit
    (the resize already yields float), alpha is a Mul, beta an Add."""
    ops: List[_Op] = []
    if alpha is not None:
        ops.append(multiply(alpha))
    if beta is not None:
        ops.append(add(beta))
    return ops


def cvtColor(code: int = COLOR_RGB2BGR):
    """cvGS::cvtColor<CODE, I, O>() :151-161 -> fk::ColorConversion<CODE> (color_conversion.cuh:364-461): the R<->B
    swaps are a VectorReorder; the other supported codes add an opaque alpha (255, AddOpaqueAlpha<float3, p8bit>), drop
    the alpha or reduce to gray, each optionally behind the swap.  Returns one op or a list of two."""
    swap3, swap4 = _Op(_abi.OP_REORDER, perm=(2, 1, 0, 3)), _Op(_abi.OP_REORDER, perm=(2, 1, 0, 3))
    alpha = _Op(_abi.OP_ADD_ALPHA, v=(255.0, 0.0, 0.0, 0.0))
    drop = _Op(_abi.OP_DROP_ALPHA)
    # the stand-alone FMUL of the gray formula differs between the instantiations (include/cvgs_b200.h, CVGS_OP_GRAY)
    gray_x, gray_y = _Op(_abi.OP_GRAY, perm=(0, 0, 0, 0)), _Op(_abi.OP_GRAY, perm=(1, 0, 0, 0))
    table = {COLOR_RGB2BGR: swap3, COLOR_RGBA2BGRA: swap4, COLOR_BGR2BGRA: alpha, COLOR_BGRA2BGR: drop,
             COLOR_BGR2RGBA: [swap3, alpha], COLOR_RGBA2BGR: [swap4, drop], COLOR_BGR2GRAY: [swap3, gray_x],
             COLOR_RGB2GRAY: gray_y, COLOR_BGRA2GRAY: [swap4, gray_x], COLOR_RGBA2GRAY: gray_y}
    if code not in table:
        raise CvgsError("colour conversion code not supported (reference cv2cuda_types.cuh:77-86 lists the set)")
    return table[code]


def split(out, planeDims: Optional[Tuple[int, int]] = None, plane_stride: int = 0) -> _Write:
    """cvGS::split<CV_32FC3>(GpuMat out, Size plane) :185-192 -> fk::TensorSplit (NCHW)."""
    return _Write(out.data_ptr(), OUT_NCHW, plane_stride, out)


def split_planes(planes) -> _Write:
    """cvGS::split<CV_32FC3>(vector<GpuMat>) / split<CV_32FC3, N>(array<vector<GpuMat>, N>) :163-183 -> fk::SplitWrite:
    `planes` = per crop three 2-D float32 CUDA tensors (rows contiguous, any row pitch), or one such triple."""
    if len(planes) == 3 and not isinstance(planes[0], (list, tuple)):
        planes = [planes]
    flat = [t for crop in planes for t in crop]
    if any(len(crop) != 3 for crop in planes):
        raise CvgsError("three destination images per crop are required")
    arr = (_abi.Plane * max(1, len(flat)))()
    for i, t in enumerate(flat):
        if t.dim() != 2 or t.stride(1) != 1:
            raise CvgsError("destination images must be 2-D with contiguous rows")
        arr[i].data, arr[i].pitch_bytes = t.data_ptr(), t.stride(0) * 4
    return _Write(C.addressof(arr), _abi.OUT_PLANES, 0, flat, arr)


def splitT(out, plane_stride: int = 0) -> _Write:
    """cvGS::splitT<CV_32FC3>(RawPtr<T3D>) :199-202 -> fk::TensorTSplit (CNHW)."""
    return _Write(out.data_ptr(), OUT_CNHW, plane_stride, out)


def write(out, plane_stride: int = 0) -> _Write:
    """cvGS::write<CV_32FC3>(GpuMat, Size) :454-457 -> packed NHWC."""
    return _Write(out.data_ptr(), OUT_NHWC, plane_stride, out)


def write_u8(out, row_pitch: int = 0, plane_stride: int = 0, cast: bool = False) -> _Write:
    """cvGS::write<CV_8UC3>(GpuMat) behind convertTo<CV_32FC3, CV_8UC3> (cast=False: SaturateCast, rounding) or behind
    fk::Cast<float3, uchar3> (cast=True: truncation, reference tests/warping/test_warping_opencv.cu:63)."""
    return _Write(out.data_ptr(), OUT_NHWC, plane_stride, out, None, _abi.CVGS_8UC3, int(row_pitch), 1 if cast else 0)


def _stream_ptr(stream) -> int:
    if stream is None:
        import torch
        return torch.cuda.current_stream().cuda_stream
    if isinstance(stream, int):
        return stream
    return stream.cuda_stream


def _flatten(ops):
    for o in ops:
        if isinstance(o, (list, tuple)):
            yield from _flatten(o)
        else:
            yield o


def build_pipeline(dsize, ops: Sequence[_Op], background=(0, 0, 0), aspect=IGNORE_AR,
                   fp_contract=FP_REFERENCE_FUSED, interp_mode=INTERP_FLOAT, out_ptr=0, layout=OUT_NCHW,
                   plane_stride=0, src_type=_abi.CVGS_8UC3, yuv_standard=0) -> _abi.Pipeline:
    p = _abi.Pipeline()
    p.src_type = int(src_type)
    p.yuv_standard = int(yuv_standard)
    p.dst_width, p.dst_height = int(dsize[0]), int(dsize[1])
    p.aspect_mode, p.interp_mode, p.fp_contract = int(aspect), int(interp_mode), int(fp_contract)
    bg = _scalar3(background)
    for c in range(4):
        p.background[c] = bg[c]
    ops = list(_flatten(ops))
    if len(ops) > _abi.MAX_OPS:
        raise CvgsError(f"at most {_abi.MAX_OPS} operations between resize and write")
    p.n_ops = len(ops)
    for i, o in enumerate(ops):
        p.ops[i].kind = o.kind
        for c in range(4):
            p.ops[i].v[c] = (tuple(o.v) + (0.0,) * 4)[c]
            p.ops[i].perm[c] = (tuple(o.perm) + (3,))[c] if c < len(o.perm) + 1 else c
    p.out_layout, p.out, p.out_plane_stride = int(layout), out_ptr, int(plane_stride)
    return p


def make_crops(mats: Sequence[GpuMat]):
    arr = (_abi.Crop * max(1, len(mats)))()
    for i, m in enumerate(mats):
        arr[i].data, arr[i].width, arr[i].height, arr[i].pitch, arr[i].reserved = m.data, m.cols, m.rows, m.step, 0
    return arr


def executeOperations(stream, *iops, fp_contract: int = FP_REFERENCE_FUSED, interp_mode: int = INTERP_FLOAT,
                      replicas: Sequence[int] = ()) -> None:
    """cvGS::executeOperations(stream, resize(...), ops..., split(...)) :464-473: ONE kernel launch,
    asynchronous on `stream`.  Raises CvgsError (the reference throws std::runtime_error).
    replicas: device pointers of further tensors (same layout, e.g. other GPUs' copies mapped into this process) that
    receive the same planes from the same launch (cvgs_b200_preproc_launch_replicated)."""
    iops = list(_flatten(iops))
    if len(iops) < 2 or not isinstance(iops[0], (_Resize, _Warp)) or not isinstance(iops[-1], _Write):
        raise CvgsError("chain must start with resize(...) / warp(...) and end with split/splitT/write(...)")
    rs, wr, mid = iops[0], iops[-1], iops[1:-1]
    if any(not isinstance(o, _Op) for o in mid):
        raise CvgsError("only multiply/subtract/divide/add/convertTo/cvtColor may sit between read and write")
    lib = _abi.load()
    if isinstance(rs, _Warp):
        p = build_pipeline(rs.dsize, mid, rs.background, IGNORE_AR, fp_contract, INTERP_FLOAT, wr.out_ptr, wr.layout,
                           wr.plane_stride, rs.src_type)
        p.dst_type, p.out_row_pitch, p.u8_cast = wr.dst_type, wr.row_pitch, wr.u8_cast
        images = make_crops(rs.images[:rs.used])
        warps = (_abi.Warp * max(1, rs.used))()
        for i, (t, v) in enumerate(rs.inverse[:rs.used]):
            warps[i].type = t
            for k in range(9):
                warps[i].m[k] = float(v[k])
        _abi.check(lib.cvgs_b200_warp_launch(images, warps, len(rs.images), rs.used, C.byref(p), _stream_ptr(stream)))
        return
    p = build_pipeline(rs.dsize, mid, rs.background, rs.aspect, fp_contract, interp_mode, wr.out_ptr, wr.layout,
                       wr.plane_stride, rs.src_type, rs.yuv_standard)
    p.dst_type, p.out_row_pitch, p.u8_cast = wr.dst_type, wr.row_pitch, wr.u8_cast
    crops = make_crops(rs.crops[:rs.used])
    parents = (_abi.Parent * max(1, rs.used))()
    for i, m in enumerate(rs.crops[:rs.used]):
        parents[i].datastart, parents[i].whole_width, parents[i].whole_height = m.datastart, m.whole[0], m.whole[1]
    if rs.src_type != _abi.CVGS_8UC3 and not (_abi.CVGS_NV12 <= rs.src_type <= _abi.CVGS_Y210):
        # parents describe 3-byte pixels (the TMA-staged kernel's domain): other depths go without
        _abi.check(lib.cvgs_b200_preproc_launch(crops, len(rs.crops), rs.used, C.byref(p), _stream_ptr(stream)))
        return
    if _abi.CVGS_NV12 <= rs.src_type <= _abi.CVGS_Y210:  # whole frames: nothing to say about parents
        _abi.check(lib.cvgs_b200_preproc_launch(crops, len(rs.crops), rs.used, C.byref(p), _stream_ptr(stream)))
        return
    if replicas:
        reps = (C.c_void_p * len(replicas))(*[int(r) for r in replicas])
        _abi.check(lib.cvgs_b200_preproc_launch_replicated(crops, parents, len(rs.crops), rs.used, C.byref(p), reps,
                                                           len(replicas), _stream_ptr(stream)))
        return
    _abi.check(lib.cvgs_b200_preproc_launch_ex(crops, parents, len(rs.crops), rs.used, C.byref(p), _stream_ptr(stream)))


def device_view(ptr: int, shape, dtype_str: str = "<f4"):
    """torch view (no copy, no ownership) of device memory somebody else allocated."""
    import torch

    class _Holder:
        pass

    h = _Holder()
    h.__cuda_array_interface__ = {"shape": tuple(int(v) for v in shape), "typestr": dtype_str, "data": (int(ptr), False),
                                  "version": 2, "strides": None}
    return torch.as_tensor(h, device="cuda")


class CircularTensor:
    """cvGS::CircularTensor<CV_8UC3, CV_32F, 3, BATCH, ORDER, MODE> (reference
    include/cvGPUSpeedup.cuh:600-627): ``update(stream, frame, ops...)`` processes the new frame into
    the newest plane and shifts the others; ``data()`` is the dense time-ordered tensor."""

    def __init__(self, width: int, height: int, batch: int, order: int = CT_NEWEST_FIRST,
                 mode: int = CT_STANDARD, color_planes: int = 3, device: int = 0, elem_channels: int = 1,
                 src_type: int = _abi.CVGS_8UC3):
        """color_planes planes of floats (1, 3 or 4), or -- color_planes 1, elem_channels 3 / 4 -- one plane of packed
        float pixels (cvGS::CircularTensor<CV_8UC4, CV_32FC4, 1, ...>).  src_type: the frames' pixel type."""
        self._lib = _abi.load()
        self._h = C.c_void_p()
        self.width, self.height, self.batch, self.order, self.mode, self.color_planes = (
            width, height, batch, order, mode, color_planes)
        self.elem_channels, self.src_type = elem_channels, src_type
        _abi.check(self._lib.cvgs_b200_ct_create_ex(C.byref(self._h), width, height, color_planes, elem_channels, batch,
                                                    order, mode, device))

    def update(self, stream, frame: GpuMat, *ops, fp_contract: int = FP_REFERENCE_FUSED,
               interp_mode: int = INTERP_FLOAT) -> None:
        p = build_pipeline((self.width, self.height), list(_flatten(ops)), fp_contract=fp_contract,
                           interp_mode=interp_mode, src_type=self.src_type)
        crop = make_crops([frame])
        _abi.check(self._lib.cvgs_b200_ct_update(self._h, crop, C.byref(p), _stream_ptr(stream)))

    def data_ptr(self) -> int:
        return int(self._lib.cvgs_b200_ct_data(self._h) or 0)

    def data(self):
        """Dense tensor as a torch view (no copy) of the library-owned memory."""
        import torch
        shape = ((self.batch, self.color_planes, self.height, self.width) if self.mode == CT_STANDARD
                 else (self.color_planes, self.batch, self.height, self.width))
        if self.elem_channels > 1:
            shape = shape + (self.elem_channels,)

        class _Holder:
            pass

        h = _Holder()
        h.__cuda_array_interface__ = {"shape": shape, "typestr": "<f4", "data": (self.data_ptr(), False),
                                      "version": 2, "strides": None}
        return torch.as_tensor(h, device="cuda")

    def close(self) -> None:
        if self._h:
            self._lib.cvgs_b200_ct_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
