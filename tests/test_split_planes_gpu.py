"""fk::SplitWrite output form (cvGS::split(vector<GpuMat>) / split(array<vector<GpuMat>, N>), reference
include/cvGPUSpeedup.cuh:163-183, memory_operations.cuh:331-360; exercised by tests/resize/test_resize_x_split.cu):
every (crop, channel) goes to its own 2-D float image with its own pointer and pitch."""
import ctypes as C

import numpy as np
import pytest
import torch

from cvgpuspeedup_b200 import _abi
import cvgpuspeedup_b200 as cvGS
from tests import util

pytestmark = pytest.mark.gpu


def _oracle_planes(w, n_planes, used, pitches, **kw):
    """Oracle on host images of the same geometry; returns [n_planes][3] arrays (H x pitch floats)."""
    W, H = w.dsize
    bufs = [[np.full((H, pitches[(3 * z + c) % len(pitches)]), np.nan, dtype=np.float32) for c in range(3)] for z in range(n_planes)]
    arr = (_abi.Plane * (3 * n_planes))()
    for z in range(n_planes):
        for c in range(3):
            arr[3 * z + c].data, arr[3 * z + c].pitch_bytes = bufs[z][c].ctypes.data, bufs[z][c].strides[0]
    p = util.make_pipeline(w.dsize, w.ops, aspect=w.aspect, background=w.background, layout=_abi.OUT_PLANES,
                           out_ptr=C.addressof(arr), **kw)
    assert util.oracle_lib().oracle_preproc(util.host_crops(w.image, w.rects[:used]), n_planes, used, C.byref(p), 0) == 0
    return bufs


@pytest.mark.parametrize("variant", [1, 2])
@pytest.mark.parametrize("case", ["c2", "ragged_ar", "single"])
def test_split_planes_matches_oracle(variant, case):
    if case == "c2":
        w, n_planes, used = util.workload_c2(seed=31, n=20, frame=(640, 480), pitch=1920), 20, 20
    elif case == "single":  # tests/resize/test_resize_x_split.cu:51,72-84: crop (200,200)-(260,320) -> 64x128
        w = util.workload_c2(seed=32, n=1, frame=(640, 480), pitch=1920)
        w.rects, n_planes, used = [(200, 200, 60, 120)], 1, 1
        w.ops = util.OPS_C1
    else:
        rng = np.random.default_rng(5)
        img = util.make_image(rng, 400, 300, pitch=1280)
        rects = [(i, i, 30, 120) for i in range(5)] + [(0, 0, 400, 300), (3, 3, 300, 9)]
        w = util.Workload("ar_planes", img, 400, 300, rects, (70, 33), util.OPS_C2, aspect=_abi.PRESERVE_AR,
                          background=(128.0, 7.0, 250.5))
        n_planes, used = 9, 7
    W, H = w.dsize
    pitches = [W, W + 3, W + 8]  # floats per row: tight and padded destination images
    want = _oracle_planes(w, n_planes, used, pitches)
    d_img = torch.from_numpy(w.image).cuda()
    frame = cvGS.GpuMat(d_img.data_ptr(), w.width, w.height, w.pitch, owner=d_img)
    crops = [frame.roi(*r) for r in w.rects[:used]] + [frame] * (n_planes - used)
    planes = [[torch.full((H, pitches[(3 * z + c) % 3]), float("nan"), device="cuda")[:, :] for c in range(3)]
              for z in range(n_planes)]
    lib = _abi.load()
    prev = lib.cvgs_b200_set_kernel_variant(variant)
    try:
        ops = []
        for k, v in w.ops:
            ops.append(cvGS.cvtColor() if k == "reorder" else {"mul": cvGS.multiply, "sub": cvGS.subtract, "div": cvGS.divide,
                                                                "add": cvGS.add}[k](v))
        cvGS.executeOperations(None, cvGS.resize(crops, w.dsize, used, w.background, w.aspect), *ops,
                               cvGS.split_planes(planes))
    finally:
        lib.cvgs_b200_set_kernel_variant(prev)
    torch.cuda.synchronize()
    for z in range(n_planes):
        for c in range(3):
            got = planes[z][c].cpu().numpy()
            util.assert_bit_equal(got[:, :W], want[z][c][:, :W], f"{case} plane {z} channel {c}")
            assert np.isnan(got[:, W:]).all(), "padding of a destination image was written"


def test_split_planes_rejects_bad_destinations():
    w = util.workload_c2(seed=33, n=2, frame=(320, 240), pitch=960)
    d_img = torch.from_numpy(w.image).cuda()
    crops = util.host_crops(w.image, w.rects, base_ptr=d_img.data_ptr())
    arr = (_abi.Plane * 6)()  # NULL data pointers
    p = util.make_pipeline(w.dsize, w.ops, layout=_abi.OUT_PLANES, out_ptr=C.addressof(arr))
    with pytest.raises(_abi.CvgsError):
        _abi.check(_abi.load().cvgs_b200_preproc_launch(crops, 2, 2, C.byref(p), None))
