"""Descriptor marshalling for the C-ABI (include/cvgs_b200.h): pipelines, crop / parent arrays, frame-loop argument sets.

What a C or C++ caller writes as struct initialisers.  Used by api.py's callers that talk to the C-ABI directly
(bench.py, __graft_entry__.smoke(), the tests); numpy-free, torch-free.
"""
from __future__ import annotations

import ctypes as C
from typing import Sequence, Tuple

from . import _abi

Rect = Tuple[int, int, int, int]  # x, y, w, h

OP_KINDS = {"mul": _abi.OP_MUL, "sub": _abi.OP_SUB, "div": _abi.OP_DIV, "add": _abi.OP_ADD, "reorder": _abi.OP_REORDER,
            "add_alpha": _abi.OP_ADD_ALPHA, "drop_alpha": _abi.OP_DROP_ALPHA, "gray": _abi.OP_GRAY}


def make_pipeline(dsize, ops, aspect=_abi.IGNORE_AR, background=(0, 0, 0), fp_contract=_abi.FP_REFERENCE_FUSED,
                  interp_mode=_abi.INTERP_FLOAT, layout=_abi.OUT_NCHW, out_ptr=0, plane_stride=0,
                  src_type=_abi.CVGS_8UC3, dst_type=0, row_pitch=0, yuv_standard=0, u8_cast=0) -> _abi.Pipeline:
    """cvgs_pipeline_t from a list of (kind, values): kind in OP_KINDS, values = per-channel constants (or the
    permutation of "reorder" / the FMUL selector of "gray")."""
    p = _abi.Pipeline()
    p.src_type = src_type
    p.dst_width, p.dst_height = dsize
    p.aspect_mode, p.interp_mode, p.fp_contract = aspect, interp_mode, fp_contract
    for c in range(len(background)):
        p.background[c] = background[c]
    p.n_ops = len(ops)
    for i, (k, v) in enumerate(ops):
        p.ops[i].kind = OP_KINDS[k]
        for c in range(len(v)):
            if k in ("reorder", "gray"):
                p.ops[i].perm[c] = v[c]
            else:
                p.ops[i].v[c] = v[c]
    p.out_layout, p.out, p.out_plane_stride = layout, out_ptr, plane_stride
    p.dst_type, p.out_row_pitch = dst_type, row_pitch
    p.yuv_standard = yuv_standard
    p.u8_cast = u8_cast
    return p


def crop_array(base_ptr: int, pitch: int, rects: Sequence[Rect], px_bytes: int = 3):
    """cvgs_crop_t[]: ROIs of the image at base_ptr (row pitch in bytes); px_bytes = 3 for CV_8UC3."""
    arr = (_abi.Crop * max(1, len(rects)))()
    for i, (x, y, w, h) in enumerate(rects):
        arr[i].data, arr[i].width, arr[i].height, arr[i].pitch, arr[i].reserved = base_ptr + y * pitch + px_bytes * x, w, h, pitch, 0
    return arr


def parent_array(base_ptr: int, width: int, height: int, n: int):
    """cvgs_parent_t[]: every crop was cut from the one image (GpuMat::datastart + locateROI)."""
    arr = (_abi.Parent * max(1, n))()
    for i in range(n):
        arr[i].datastart, arr[i].whole_width, arr[i].whole_height = base_ptr, width, height
    return arr


class FrameSets:
    """Argument sets of the C-ABI frame loops: one (frame, rect list, output tensor) per set, the same pipeline.

    frames   sequence of (device image pointer, pitch, width, height, rects)
    outs     device pointers of the per-set output tensors
    The ctypes arrays stay alive with the object (the C-ABI reads them during the call only)."""

    def __init__(self, frames, outs, dsize, ops, **pipe_kw):
        n = len(frames)
        self.n = n
        self.crop_sets = [crop_array(ptr, pitch, rects) for (ptr, pitch, _w, _h, rects) in frames]
        self.parent_sets = [parent_array(ptr, w, h, len(rects)) for (ptr, _pitch, w, h, rects) in frames]
        self.pipes = [make_pipeline(dsize, ops, out_ptr=o, **pipe_kw) for o in outs]
        self.crops_pp = (C.POINTER(_abi.Crop) * n)(*[C.cast(c, C.POINTER(_abi.Crop)) for c in self.crop_sets])
        self.parents_pp = (C.POINTER(_abi.Parent) * n)(*[C.cast(c, C.POINTER(_abi.Parent)) for c in self.parent_sets])
        self.pipes_pp = (C.POINTER(_abi.Pipeline) * n)(*[C.pointer(p) for p in self.pipes])
        self.n_arr = (C.c_int32 * n)(*[len(f[4]) for f in frames])

    def launch_sequence(self, lib, steps: int, stream_ptr: int) -> None:
        """cvgs_b200_preproc_launch_sequence_ex: `steps` frames, frame i using set i % n."""
        _abi.check(lib.cvgs_b200_preproc_launch_sequence_ex(self.crops_pp, self.parents_pp, self.n_arr, self.n_arr,
                                                            self.pipes_pp, self.n, steps, stream_ptr))
