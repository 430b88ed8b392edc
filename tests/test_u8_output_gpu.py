"""8-bit output: the chain ends with convertTo<CV_32FC3, CV_8UC3> (SaturateCast<float, uchar>, reference
saturate.cuh:127-147) and packed pixels are written with the destination's own row pitch
(cvGS::write<CV_8UC3>(GpuMat); reference tests/resize/test_resize_write.cu:55-56)."""
import ctypes as C

import numpy as np
import pytest
import torch

from cvgpuspeedup_b200 import _abi
from tests import util

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("src_type,px", [(_abi.CVGS_8UC3, 3), (_abi.CVGS_16UC3, 6)])
@pytest.mark.parametrize("aspect", [_abi.IGNORE_AR, _abi.PRESERVE_AR_LEFT])
def test_u8_output_matches_oracle(src_type, px, aspect):
    rng = np.random.default_rng(61)
    img = rng.integers(0, 256, size=(240, 2048), dtype=np.uint8)
    rects = [(0, 0, 320, 240), (3, 5, 40, 80), (100, 20, 200, 200), (319, 0, 1, 240), (7, 7, 64, 128)]
    lib = _abi.load()
    d_img = torch.from_numpy(img).cuda()
    for (W, H), pitch in [((64, 128), 0), ((33, 7), 128), ((200, 150), 608)]:
        # chain chosen so that values leave [0, 255] on both sides: the saturation is exercised
        ops = [("reorder", (2, 1, 0)), ("mul", (1.7, 0.004 if px == 6 else 1.0, -0.5)), ("sub", (40.0, -3.0, 0.25))]
        rp = pitch or 3 * W
        n = len(rects) + 1  # one unused plane: saturate(chain(background))
        want = np.full((n, H, rp), 7, dtype=np.uint8)
        got = torch.full((n, H, rp), 7, dtype=torch.uint8, device="cuda")
        kw = dict(aspect=aspect, background=(300.0, 12.6, -4.0), layout=_abi.OUT_NHWC, src_type=src_type,
                  dst_type=_abi.CVGS_8UC3, row_pitch=pitch)
        p = util.make_pipeline((W, H), ops, out_ptr=want.ctypes.data, **kw)
        assert util.oracle_lib().oracle_preproc(util.host_crops(img, rects, px_bytes=px), n, len(rects), C.byref(p), 0) == 0
        p = util.make_pipeline((W, H), ops, out_ptr=got.data_ptr(), **kw)
        _abi.check(lib.cvgs_b200_preproc_launch(util.host_crops(img, rects, base_ptr=d_img.data_ptr(), px_bytes=px), n,
                                                len(rects), C.byref(p), None))
        torch.cuda.synchronize()
        assert np.array_equal(got.cpu().numpy(), want), f"{(W, H)} pitch {pitch}"
        assert (want[:, :, :3 * W] == 255).any() and (want[:, :, :3 * W] == 0).any()


def test_u8_output_argument_checks():
    lib = _abi.load()
    w = util.workload_c2(seed=62, n=2, frame=(320, 240), pitch=960)
    d_img = torch.from_numpy(w.image).cuda()
    out = torch.zeros(2 * 128 * 64 * 3, dtype=torch.uint8, device="cuda")
    crops = util.host_crops(w.image, w.rects, base_ptr=d_img.data_ptr())
    for kw in [dict(layout=_abi.OUT_NCHW), dict(layout=_abi.OUT_NHWC, row_pitch=100)]:
        p = util.make_pipeline(w.dsize, w.ops, out_ptr=out.data_ptr(), dst_type=_abi.CVGS_8UC3, **kw)
        with pytest.raises(_abi.CvgsError):
            _abi.check(lib.cvgs_b200_preproc_launch(crops, 2, 2, C.byref(p), None))


@pytest.mark.parametrize("src_type,px", [(_abi.CVGS_8UC4, 4), (_abi.CVGS_16SC4, 8)])
def test_u8_output_4_channels(src_type, px):
    """resize<CV_8UC4> -> convertTo<CV_32FC4, CV_8UC4> -> write<CV_8UC4> (reference tests/resize/
    test_resize_CPUvsGPUresults.cu:45-47), plus a chain that saturates on both sides."""
    rng = np.random.default_rng(63)
    img = rng.integers(0, 256, size=(240, 320 * px + 64), dtype=np.uint8)
    rects = [(0, 0, 320, 240), (3, 5, 40, 80), (100, 20, 200, 200), (319, 0, 1, 240)]
    lib = _abi.load()
    d_img = torch.from_numpy(img).cuda()
    for ops in ([], [("reorder", (2, 1, 0, 3)), ("mul", (1.7, 0.004 if px == 8 else 1.0, -0.5, 1.0)), ("sub", (40.0, -3.0, 0.25, 0.5))]):
        for (W, H), pitch in [((64, 128), 0), ((33, 7), 160)]:
            rp = pitch or 4 * W
            n = len(rects) + 1
            want = np.full((n, H, rp), 7, dtype=np.uint8)
            got = torch.full((n, H, rp), 7, dtype=torch.uint8, device="cuda")
            kw = dict(background=(300.0, 12.6, -4.0, 77.5), layout=_abi.OUT_NHWC, src_type=src_type, dst_type=_abi.CVGS_8UC4,
                      row_pitch=pitch)
            p = util.make_pipeline((W, H), ops, out_ptr=want.ctypes.data, **kw)
            assert util.oracle_lib().oracle_preproc(util.host_crops(img, rects, px_bytes=px), n, len(rects), C.byref(p), 0) == 0
            p = util.make_pipeline((W, H), ops, out_ptr=got.data_ptr(), **kw)
            _abi.check(lib.cvgs_b200_preproc_launch(util.host_crops(img, rects, base_ptr=d_img.data_ptr(), px_bytes=px), n,
                                                    len(rects), C.byref(p), None))
            torch.cuda.synchronize()
            assert np.array_equal(got.cpu().numpy(), want), f"{ops} {(W, H)} pitch {pitch}"


@pytest.mark.parametrize("src_type,px,nc", [(_abi.CVGS_8UC3, 3, 3), (_abi.CVGS_8UC4, 4, 4), (_abi.CVGS_16UC3, 6, 3)])
def test_packed_float_output_with_padded_rows(src_type, px, nc):
    """PerThreadWrite<_2D, float3 / float4> into a GpuMat whose step is larger than a row (cudaMallocPitch): the form of
    cvGS::executeOperations(input, output, stream, ops...) (reference include/cvGPUSpeedup.cuh:489-503,
    tests/read/test_read_x_write.cu).  Bytes between the rows must stay untouched."""
    rng = np.random.default_rng(64)
    img = rng.integers(0, 256, size=(120, 160 * px + 32), dtype=np.uint8)
    rects = [(0, 0, 160, 120), (10, 20, 50, 60), (3, 3, 9, 100)]
    lib = _abi.load()
    d_img = torch.from_numpy(img).cuda()
    ops = [("sub", (1.0, 4.0, 3.2, 0.5)[:nc]), ("mul", (0.3, 0.5, 2.0, 1.5)[:nc]), ("div", (3.2, 0.6, 11.8, 2.0)[:nc]),
           ("add", (0.5, 1.5, 2.5, 3.5)[:nc])]
    for (W, H), pitch_floats in [((160, 120), 160 * nc + 16), ((33, 7), 33 * nc + 5), ((64, 48), 64 * nc)]:
        n = len(rects)
        want = np.full((n, H, pitch_floats), -7.0, dtype=np.float32)
        got = torch.full((n, H, pitch_floats), -7.0, dtype=torch.float32, device="cuda")
        kw = dict(layout=_abi.OUT_NHWC, src_type=src_type, row_pitch=4 * pitch_floats)
        p = util.make_pipeline((W, H), ops, out_ptr=want.ctypes.data, **kw)
        assert util.oracle_lib().oracle_preproc(util.host_crops(img, rects, px_bytes=px), n, n, C.byref(p), 0) == 0
        p = util.make_pipeline((W, H), ops, out_ptr=got.data_ptr(), **kw)
        par = util.host_parents(img, 160, 120, n, base_ptr=d_img.data_ptr())
        _abi.check(lib.cvgs_b200_preproc_launch_ex(util.host_crops(img, rects, base_ptr=d_img.data_ptr(), px_bytes=px), par, n, n,
                                                   C.byref(p), None))
        torch.cuda.synchronize()
        util.assert_bit_equal(got.cpu().numpy(), want, f"src {src_type} {(W, H)} pitch {pitch_floats}")
        if pitch_floats > W * nc:
            assert (want[:, :, W * nc:] == -7.0).all()


def test_u8_output_through_the_tma_kernel():
    """CV_8UC3 -> CV_8UC3 with 16-byte aligned pitches is taken by the TMA-staged kernel (general instantiation): forced
    variant 2 must accept it and agree with the oracle, for both casts, with padded rows, unused planes and AR bands."""
    rng = np.random.default_rng(65)
    w = util.workload_c2(seed=66, n=40, frame=(640, 480), pitch=1920)
    lib = _abi.load()
    d_img = torch.from_numpy(w.image).cuda()
    prev = lib.cvgs_b200_set_kernel_variant(2)
    try:
        for (W, H), pitch, cast, aspect, ops in [((64, 128), 0, 0, _abi.IGNORE_AR, []), ((100, 60), 320, 0, _abi.PRESERVE_AR, util.OPS_C2),
                                                 ((33, 7), 0, 1, _abi.IGNORE_AR, []), ((224, 224), 700, 0, _abi.PRESERVE_AR_LEFT,
                                                                                     [("reorder", (2, 1, 0)), ("mul", (1.7, 1.0, -0.5)), ("sub", (40.0, -3.0, 0.25))])]:
            rp = pitch or 3 * W
            n = len(w.rects) + 2
            want = np.full((n, H, rp), 7, dtype=np.uint8)
            got = torch.full((n, H, rp), 7, dtype=torch.uint8, device="cuda")
            kw = dict(aspect=aspect, background=(300.0, 12.6, -4.0), layout=_abi.OUT_NHWC, dst_type=_abi.CVGS_8UC3, row_pitch=pitch,
                      u8_cast=cast)
            if cast:
                kw["background"] = (30.0, 12.6, 4.0)  # fk::Cast is defined for values inside [0, 256)
            p = util.make_pipeline((W, H), ops, out_ptr=want.ctypes.data, **kw)
            assert util.oracle_lib().oracle_preproc(util.host_crops(w.image, w.rects), n, len(w.rects), C.byref(p), 0) == 0
            p = util.make_pipeline((W, H), ops, out_ptr=got.data_ptr(), **kw)
            par = util.host_parents(w.image, w.width, w.height, len(w.rects), base_ptr=d_img.data_ptr())
            _abi.check(lib.cvgs_b200_preproc_launch_ex(util.host_crops(w.image, w.rects, base_ptr=d_img.data_ptr()), par, n,
                                                       len(w.rects), C.byref(p), None))
            torch.cuda.synchronize()
            assert np.array_equal(got.cpu().numpy(), want), f"{(W, H)} pitch {pitch} cast {cast}"
    finally:
        lib.cvgs_b200_set_kernel_variant(prev)
