"""4-channel sources (CV_8UC4 / CV_16UC4 / CV_16SC4 -> CV_32FC4: the other half of the reference's type matrix,
tests/batchresize/test_batchresize_x_split3D.cu:427-432, chain with cvtColor<COLOR_RGBA2BGRA>): bit-exact against the
oracle and against the reference's own uchar4 / ushort4 / short4 instantiations."""
import ctypes as C

import numpy as np
import pytest
import torch

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


def test_tma_kernel_declines_other_geometries_of_16bit_sources_and_unaligned_8uc4():
    """The TMA-staged kernel takes CV_8UC4 whose pixels are aligned words and 16-bit pixels in the common geometry; the
    aspect-ratio modes of 16-bit sources and 8UC4 images at odd byte offsets belong to the direct-gather kernel (variant 2 =
    fail instead of falling back)."""
    img = _img(90, _abi.CVGS_16SC4)
    with pytest.raises(_abi.CvgsError):
        gpu_util.run_cvgs(img, [(0, 0, 64, 64)], (32, 32), [], variant=2, src_type=_abi.CVGS_16SC4, aspect=_abi.PRESERVE_AR)
    lib = _abi.load()
    buf = torch.zeros(64 * 272 + 8, dtype=torch.uint8, device="cuda")
    out = torch.empty((1, 4, 32, 32), device="cuda")
    crop = (_abi.Crop * 1)()
    crop[0].data, crop[0].width, crop[0].height, crop[0].pitch = buf.data_ptr() + 2, 64, 64, 272
    p = util.make_pipeline((32, 32), [], out_ptr=out.data_ptr(), src_type=_abi.CVGS_8UC4)
    prev = lib.cvgs_b200_set_kernel_variant(2)
    try:
        assert lib.cvgs_b200_preproc_launch(crop, 1, 1, C.byref(p), None) == 801
    finally:
        lib.cvgs_b200_set_kernel_variant(prev)


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


@pytest.mark.parametrize("src_type,px", [(_abi.CVGS_8UC4, 4), (_abi.CVGS_16UC3, 6), (_abi.CVGS_16UC4, 8), (_abi.CVGS_16SC3, 6), (_abi.CVGS_16SC4, 8)])
@pytest.mark.parametrize("n,parents", [(8, True), (100, True), (300, True), (40, False), (300, False)])
def test_tma_staged_kernel_takes_8uc4_and_16u(n, parents, src_type, px):
    """CV_8UC4 / CV_16UC3 / CV_16UC4 / CV_16SC3 / CV_16SC4 through the TMA-staged kernel (forced: variant 2 fails instead of
    falling back; signed samples: sign bit flipped, converted as unsigned, 2^-126 subtracted in the scaled domain):
    aligned-word pixels, halfword samples (6-byte pixels at both word phases), three / four channels through chain and
    stores.  Small batches, the 256-crop table and the descriptor ring; with and without parent images; mixed up- and
    down-scales; the FMA-DIV chain, a generic chain, no chain, and the rounded-resize mode."""
    lib = _abi.load()
    rng = np.random.default_rng(60 + n + px)
    fw, fh = 640, 480
    pitch = px * fw + 64
    img = rng.integers(0, 256, size=(fh, pitch), dtype=np.uint8)
    nc = util.channels_of(src_type)
    rects = [(1, 0, 33, 20), (2, 1, 17, 9)]  # odd x: 6-byte pixels on the other word phase
    top = 300 if px < 8 else 200  # 8-byte pixels: a 32-column band of a 9x down-scale would exceed the 2 KB TMA box row
    for _ in range(n - 2):
        w, h = int(rng.integers(8, top)), int(rng.integers(8, top))
        rects.append((int(rng.integers(0, fw - w + 1)), int(rng.integers(0, fh - h + 1)), w, h))
    d_img = torch.from_numpy(img).cuda()
    cases = [((64, 128), [("reorder", (2, 1, 0, 3)[:nc]), ("mul", (0.3, 0.3, 0.3, 0.5)[:nc]), ("sub", (1.0, 4.0, 3.2, 0.25)[:nc]),
                          ("div", (3.2, 0.6, 11.8, 2.0)[:nc])], {}),
             ((224, 224), [("mul", (0.5, 0.25, 2.0, 1.0)[:nc]), ("add", (1.0, 2.0, 3.0, 4.0)[:nc])], {}),
             ((33, 17), [], {}),
             ((48, 40), [("mul", (0.5, 0.25, 2.0, 1.0)[:nc])], dict(fp_contract=_abi.FP_SEPARATE, interp_mode=_abi.INTERP_ROUND_U8))]
    for dsize, ops, kw in cases:
        out = torch.full((n, nc, dsize[1], dsize[0]), float("nan"), device="cuda")
        p = util.make_pipeline(dsize, ops, out_ptr=out.data_ptr(), src_type=src_type, **kw)
        crops = util.host_crops(img, rects, base_ptr=d_img.data_ptr(), px_bytes=px)
        par = util.host_parents(img, fw, fh, n, base_ptr=d_img.data_ptr()) if parents else None
        prev = lib.cvgs_b200_set_kernel_variant(2)
        try:
            _abi.check(lib.cvgs_b200_preproc_launch_ex(crops, par, n, n, C.byref(p), None))
        finally:
            lib.cvgs_b200_set_kernel_variant(prev)
        torch.cuda.synchronize()
        util.assert_bit_equal(out.cpu().numpy(), util.run_oracle(img, rects, dsize, ops, src_type=src_type, **kw), f"{src_type} {dsize}")
