"""cvgs_b200_preproc_launch_ex: crops staged through cached per-image tensor maps (the caller names the parent image,
like cv::cuda::GpuMat::datastart / locateROI) must give exactly the results of the per-crop path and of the oracle."""
import ctypes as C

import numpy as np
import pytest
import torch

from cvgpuspeedup_b200 import _abi
from tests import gpu_util, util

pytestmark = pytest.mark.gpu


def _check(w: util.Workload, **kw):
    got = gpu_util.run_cvgs(w.image, w.rects, w.dsize, w.ops, variant=2, aspect=w.aspect, background=w.background,
                            parents=(w.width, w.height), **kw)
    kw.pop("d_image", None)
    want = util.run_oracle(w.image, w.rects, w.dsize, w.ops, aspect=w.aspect, background=w.background, **kw)
    util.assert_bit_equal(got, want, f"{w.name} image maps {kw}")


def test_baseline_configs_through_image_maps():
    _check(util.workload_c1())
    _check(util.workload_c2(pitch=6144))                    # mixed scales: several row-byte classes of one image
    _check(util.workload_c2(pitch=5760, ref_shape=True))    # tight pitch, crops (i, i, 60, 120)
    _check(util.workload_c3(n=40))                          # two column bands per crop
    _check(util.workload_c3(n=200))                         # > 64 crops: descriptors through the device ring


@pytest.mark.parametrize("dsize", [(63, 17), (1, 1), (5, 300), (130, 2), (224, 224), (300, 40)])
def test_edges_of_the_parent_image(dsize):
    """Crops that touch every border of the parent: the staging boxes overhang the crop but never the image."""
    rng = np.random.default_rng(17)
    img = util.make_image(rng, 333, 211, pitch=1008)
    rects = [(0, 0, 333, 211), (1, 1, 1, 1), (332, 210, 1, 1), (0, 0, 2, 2), (7, 9, 100, 3), (300, 5, 33, 200),
             (10, 10, 64, 128), (11, 12, 65, 129), (5, 0, 328, 1), (0, 210, 333, 1), (332, 0, 1, 211)]
    w = util.Workload("edges_img", img, 333, 211, rects, dsize, util.OPS_C1)
    _check(w)
    _check(w, fp_contract=_abi.FP_SEPARATE, interp_mode=_abi.INTERP_ROUND_U8)


@pytest.mark.parametrize("aspect", [_abi.PRESERVE_AR, _abi.PRESERVE_AR_RN_EVEN, _abi.PRESERVE_AR_LEFT])
def test_aspect_modes_layouts_unused_planes(aspect):
    rng = np.random.default_rng(5)
    img = util.make_image(rng, 400, 300, pitch=1280)
    rects = [(i, i, 30, 120) for i in range(8)] + [(0, 0, 400, 300), (10, 10, 17, 200), (3, 3, 300, 9)]
    w = util.Workload("ar_img", img, 400, 300, rects, (64, 128), util.OPS_C2, aspect=aspect, background=(128.0, 7.0, 250.5))
    _check(w)
    _check(w, layout=_abi.OUT_NHWC, n_planes=14, used=9)
    _check(w, layout=_abi.OUT_CNHW, plane_stride=64 * 128 + 4)


def test_unaligned_parent_base():
    rng = np.random.default_rng(23)
    back = rng.integers(0, 256, size=240 * 960 + 64, dtype=np.uint8)
    d_back = torch.from_numpy(back).cuda()
    rects = [(0, 0, 319, 240), (3, 5, 40, 80), (100, 20, 200, 200), (318, 0, 1, 240), (7, 7, 64, 128)]
    for shift in [0, 1, 5, 8, 15]:
        img = back[shift:shift + 240 * 960].reshape(240, 960)
        d_img = d_back[shift:shift + 240 * 960].view(240, 960)
        w = util.Workload("shift_img", img, 319, 240, rects, (64, 128), util.OPS_C2)
        _check(w, d_image=d_img)


def test_crops_of_several_images_and_fallbacks():
    """Two parent images in one launch; a parent that does not contain its crop and a NULL parent fall back to the
    per-crop path for the whole launch -- same results."""
    lib = _abi.load()
    rng = np.random.default_rng(3)
    imgs = [util.make_image(rng, 320, 240, pitch=960) for _ in range(2)]
    d_imgs = [torch.from_numpy(i).cuda() for i in imgs]
    rects = [(5, 5, 100, 200), (0, 0, 320, 240), (200, 100, 64, 128), (310, 230, 10, 10)]
    dsize, ops = (64, 128), util.OPS_C2
    want = np.concatenate([util.run_oracle(imgs[k], rects, dsize, ops) for k in range(2)])
    n = 2 * len(rects)
    crops = (_abi.Crop * n)()
    for k in range(2):
        for i, (x, y, w, h) in enumerate(rects):
            c = crops[k * len(rects) + i]
            c.data, c.width, c.height, c.pitch, c.reserved = d_imgs[k].data_ptr() + y * 960 + 3 * x, w, h, 960, 0

    def run(parents):
        out = torch.full((n, 3, dsize[1], dsize[0]), float("nan"), device="cuda")
        p = util.make_pipeline(dsize, ops, out_ptr=out.data_ptr())
        prev = lib.cvgs_b200_set_kernel_variant(2)
        try:
            _abi.check(lib.cvgs_b200_preproc_launch_ex(crops, parents, n, n, C.byref(p), torch.cuda.current_stream().cuda_stream))
        finally:
            lib.cvgs_b200_set_kernel_variant(prev)
        torch.cuda.synchronize()
        return out.cpu().numpy()

    good = (_abi.Parent * n)()
    for k in range(2):
        for i in range(len(rects)):
            q = good[k * len(rects) + i]
            q.datastart, q.whole_width, q.whole_height = d_imgs[k].data_ptr(), 320, 240
    util.assert_bit_equal(run(good), want, "two parent images")
    bad = (_abi.Parent * n)(*good)
    bad[3].whole_height = 100          # crop 3 (rows 230..239) is outside this "parent"
    util.assert_bit_equal(run(bad), want, "inconsistent parent -> per-crop maps")
    bad = (_abi.Parent * n)(*good)
    bad[0].datastart = None
    util.assert_bit_equal(run(bad), want, "NULL parent -> per-crop maps")
    util.assert_bit_equal(run(None), want, "no parents")


def test_image_map_cache_survives_buffer_reuse():
    """Same device buffer, new contents and new crops on every launch (a camera ring): the cached map is keyed by
    the image, not by its contents."""
    for seed in range(6):
        w = util.workload_c2(seed=100 + seed, n=20, frame=(640, 480), pitch=1920)
        if seed == 0:
            d_img = torch.from_numpy(w.image).cuda()
        else:
            d_img.copy_(torch.from_numpy(w.image))
        _check(w, d_image=d_img)
