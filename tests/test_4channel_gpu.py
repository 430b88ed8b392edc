"""4-channel sources (CV_8UC4 / CV_16UC4 / CV_16SC4 -> CV_32FC4: the other half of the reference's type matrix,
tests/batchresize/test_batchresize_x_split3D.cu:427-432, chain with cvtColor<COLOR_RGBA2BGRA>): bit-exact against the
oracle and against the reference's own uchar4 / ushort4 / short4 instantiations."""
import numpy as np
import pytest

from cvgpuspeedup_b200 import _abi
from tests import gpu_util, util

pytestmark = pytest.mark.gpu

MUL, SUB, DIV, BG = (0.3, 0.3, 0.3, 0.25), (1.0, 4.0, 3.2, 0.5), (3.2, 0.6, 11.8, 2.0), (128.0, 3.5, 250.0, 7.0)
TYPES = [_abi.CVGS_8UC4, _abi.CVGS_16UC4, _abi.CVGS_16SC4]


def _img(seed, src_type):
    rng = np.random.default_rng(seed)
    return rng.integers(0, 256, size=(240, 320 * util.px_bytes_of(src_type) + 64), dtype=np.uint8)


RECTS = [(0, 0, 320, 240), (5, 7, 24, 48), (100, 3, 199, 33), (17, 150, 7, 5), (300, 0, 20, 240), (1, 1, 64, 128),
         (319, 239, 1, 1), (20, 20, 60, 120)]


@pytest.mark.parametrize("src_type", TYPES)
@pytest.mark.parametrize("aspect", [_abi.IGNORE_AR, _abi.PRESERVE_AR])
def test_reference_kernel_4channel(src_type, aspect):
    if gpu_util.fkref_lib(16) is None:
        pytest.skip("oracle/_ref not built")
    img = _img(70 + src_type, src_type)
    ops = [("reorder", (2, 1, 0, 3)), ("mul", MUL), ("sub", SUB), ("div", DIV)]
    ref = gpu_util.run_fkref(img, RECTS, (64, 128), 1, MUL, SUB, DIV, aspect=aspect, bg=BG, batch=16, used=8, src_type=src_type)
    ours = gpu_util.run_cvgs(img, RECTS, (64, 128), ops, n_planes=16, used=8, aspect=aspect, background=BG, src_type=src_type)
    util.assert_bit_equal(ours, ref, "4-channel: ours vs reference kernel")
    orc = util.run_oracle(img, RECTS, (64, 128), ops, n_planes=16, used=8, aspect=aspect, background=BG, src_type=src_type)
    util.assert_bit_equal(orc, ref, "4-channel: oracle vs reference kernel")


@pytest.mark.parametrize("src_type", TYPES)
def test_4channel_layouts_and_modes(src_type):
    img = _img(80 + src_type, src_type)
    ops = [("mul", MUL), ("reorder", (3, 0, 2, 1)), ("add", (0.5, 1.5, 2.5, 3.5)), ("div", DIV), ("reorder", (1, 0, 3, 2))]
    for kw in [dict(layout=_abi.OUT_NHWC), dict(layout=_abi.OUT_CNHW, plane_stride=33 * 7 + 3),
               dict(fp_contract=_abi.FP_SEPARATE, interp_mode=_abi.INTERP_ROUND_U8), dict(n_planes=10, used=6)]:
        got = gpu_util.run_cvgs(img, RECTS, (33, 7), ops, aspect=_abi.PRESERVE_AR_LEFT, background=BG, src_type=src_type, **kw)
        want = util.run_oracle(img, RECTS, (33, 7), ops, aspect=_abi.PRESERVE_AR_LEFT, background=BG, src_type=src_type, **kw)
        util.assert_bit_equal(got, want, f"src_type {src_type} {kw}")


def test_tma_kernel_declines_4channel_sources():
    img = _img(90, _abi.CVGS_8UC4)
    with pytest.raises(_abi.CvgsError):
        gpu_util.run_cvgs(img, [(0, 0, 64, 64)], (32, 32), [], variant=2, src_type=_abi.CVGS_8UC4)


@pytest.mark.parametrize("src_type,shifts", [(_abi.CVGS_8UC4, (0, 1, 2, 3)), (_abi.CVGS_16UC4, (0, 2, 4, 6)), (_abi.CVGS_16SC4, (0, 2))])
def test_4channel_taps_at_every_base_alignment(src_type, shifts):
    """Aligned images fetch a tap with one 32-/64-bit load; any other base or pitch must fall back to element loads and
    give the same bits."""
    import ctypes as C
    import torch
    px = util.px_bytes_of(src_type)
    rng = np.random.default_rng(85 + src_type)
    w, h = 100, 60
    lib = _abi.load()
    for shift in shifts:
        for pitch in (px * w + 16, px * w + 2 * (px // 4) + (2 if px == 8 else 1)):
            if px == 8 and pitch % 2:
                pitch += 1
            back = rng.integers(0, 256, size=h * pitch + 16, dtype=np.uint8)
            img = back[shift:shift + h * pitch].reshape(h, pitch)
            d_back = torch.from_numpy(back).cuda()
            rects = [(0, 0, w, h), (3, 5, 40, 30), (99, 0, 1, 60)]
            ops = [("mul", MUL), ("sub", SUB)]
            want = util.run_oracle(np.ascontiguousarray(img), rects, (37, 23), ops, src_type=src_type)
            out = torch.full((3, 4, 23, 37), float("nan"), device="cuda")
            crops = util.host_crops(np.ascontiguousarray(img), rects, base_ptr=d_back.data_ptr() + shift, px_bytes=px)
            for i in range(3):
                crops[i].pitch = pitch
            p = util.make_pipeline((37, 23), ops, out_ptr=out.data_ptr(), src_type=src_type)
            _abi.check(lib.cvgs_b200_preproc_launch(crops, 3, 3, C.byref(p), None))
            torch.cuda.synchronize()
            util.assert_bit_equal(out.cpu().numpy(), want, f"src {src_type} shift {shift} pitch {pitch}")
