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


def test_tma_kernel_takes_signed_16bit_sources_bit_exact_with_the_reference_kernel():
    """Signed 16-bit samples through the TMA-staged kernel (forced: variant 2): extreme values (-32768, 32767, 0, -1) next to
    each other, against the oracle and -- where built -- the reference's own short3 instantiation."""
    w, h, pitch = 96, 64, 6 * 96 + 16
    vals = np.array([-32768, 32767, 0, -1, 1, -12345, 12345, 255, -256], dtype=np.int16)
    rng = np.random.default_rng(43)
    img16 = vals[rng.integers(0, len(vals), size=(h, pitch // 2))]
    img = img16.view(np.uint8).reshape(h, pitch)
    rects = [(0, 0, 96, 64), (3, 5, 40, 50), (95, 0, 1, 64), (10, 10, 64, 32)]
    mul, sub, div = (1 / 257.0,) * 3, (1.0, 4.0, 3.2), (3.2, 0.6, 11.8)
    ops = [("reorder", (2, 1, 0)), ("mul", mul), ("sub", sub), ("div", div)]
    for dsize in [(64, 128), (33, 7), (200, 100)]:
        got = gpu_util.run_cvgs(img, rects, dsize, ops, variant=2, src_type=_abi.CVGS_16SC3)
        want = util.run_oracle(img, rects, dsize, ops, src_type=_abi.CVGS_16SC3)
        util.assert_bit_equal(got, want, f"signed 16-bit, TMA-staged kernel vs oracle {dsize}")
    if gpu_util.fkref_lib(16) is not None:
        rects16 = rects * 4
        ref = gpu_util.run_fkref(img, rects16, (64, 128), 1, mul, sub, div, batch=16, used=16, src_type=_abi.CVGS_16SC3)
        got = gpu_util.run_cvgs(img, rects16, (64, 128), ops, variant=2, src_type=_abi.CVGS_16SC3)
        util.assert_bit_equal(got, ref, "signed 16-bit, TMA-staged kernel vs reference kernel")


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
