"""CV_16UC3 / CV_16SC3 sources (reference test matrix tests/batchresize/test_batchresize_x_split3D.cu:427-432;
SaturateCast saturate.cuh:267-298,358-378): same arithmetic on ushort3 / short3 taps, taken by the direct-gather
kernel; the TMA-staged kernel declines them."""
import numpy as np
import pytest
import torch

from cvgpuspeedup_b200 import _abi
import cvgpuspeedup_b200 as cvGS
from tests import gpu_util, util

pytestmark = pytest.mark.gpu


def _image16(seed, width, height, pitch, signed):
    """[H, pitch] uint8 backing store whose first 6*width bytes per row are random 16-bit pixels (full range)."""
    rng = np.random.default_rng(seed)
    img = rng.integers(0, 256, size=(height, pitch), dtype=np.uint8)
    return img


@pytest.mark.parametrize("src_type", [_abi.CVGS_16UC3, _abi.CVGS_16SC3])
@pytest.mark.parametrize("aspect", [_abi.IGNORE_AR, _abi.PRESERVE_AR])
def test_16bit_sources_match_oracle(src_type, aspect):
    img = _image16(41, 320, 240, 1936, src_type == _abi.CVGS_16SC3)
    rects = [(0, 0, 320, 240), (3, 5, 40, 80), (100, 20, 200, 200), (319, 0, 1, 240), (7, 7, 64, 128), (1, 1, 1, 1)]
    ops = [("reorder", (2, 1, 0)), ("mul", (1 / 257.0,) * 3), ("sub", (1.0, 4.0, 3.2)), ("div", (3.2, 0.6, 11.8))]
    for dsize in [(64, 128), (33, 7)]:
        for kw in [dict(), dict(fp_contract=_abi.FP_SEPARATE, interp_mode=_abi.INTERP_ROUND_U8),
                   dict(layout=_abi.OUT_NHWC, n_planes=8, used=5)]:
            got = gpu_util.run_cvgs(img, rects, dsize, ops, variant=0, aspect=aspect, background=(9.5, 1e4, -3.0),
                                    src_type=src_type, **kw)
            want = util.run_oracle(img, rects, dsize, ops, aspect=aspect, background=(9.5, 1e4, -3.0), src_type=src_type, **kw)
            util.assert_bit_equal(got, want, f"src_type {src_type} aspect {aspect} {dsize} {kw}")


def test_tma_kernel_declines_signed_16bit_sources():
    """Unsigned 16-bit samples go through the TMA-staged kernel (tests/test_4channel_gpu.py); signed ones keep the
    direct-gather kernel (variant 2 = fail instead of falling back)."""
    img = _image16(42, 64, 64, 384, False)
    with pytest.raises(_abi.CvgsError):
        gpu_util.run_cvgs(img, [(0, 0, 64, 64)], (32, 32), [], variant=2, src_type=_abi.CVGS_16SC3)


def test_circular_tensor_with_16bit_frames():
    import ctypes as C
    oc = util.oracle_lib()
    ct = cvGS.CircularTensor(48, 36, 3, cvGS.CT_OLDEST_FIRST, cvGS.CT_STANDARD)
    h = oc.oracle_ct_create(48, 36, 3, 3, _abi.CT_OLDEST_FIRST, _abi.CT_STANDARD)
    lib = _abi.load()
    for i in range(4):
        img = _image16(50 + i, 96, 72, 576, False)
        d = torch.from_numpy(img).cuda()
        p = util.make_pipeline((48, 36), [("mul", (0.001,) * 3), ("sub", (1.0, 2.0, 3.0))], src_type=_abi.CVGS_16UC3)
        crop = util.host_crops(img, [(0, 0, 96, 72)], base_ptr=d.data_ptr(), px_bytes=6)
        _abi.check(lib.cvgs_b200_ct_update(ct._h, crop, C.byref(p), None))
        torch.cuda.synchronize()
        assert oc.oracle_ct_update(h, util.host_crops(img, [(0, 0, 96, 72)], px_bytes=6), C.byref(p), 0) == 0
    want = np.ctypeslib.as_array(oc.oracle_ct_data(h), shape=(3, 3, 36, 48)).copy()
    util.assert_bit_equal(ct.data().cpu().numpy(), want, "CircularTensor with CV_16UC3 frames")
    oc.oracle_ct_destroy(h)
    ct.close()
