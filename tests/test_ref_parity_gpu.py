"""The strongest parity pin available: the reference's OWN fused kernel (FusedKernelLibrary headers compiled
unmodified into oracle/_ref/libfkref_N.so by oracle/Makefile) runs next to ours on the same B200 and the
outputs must be bit-identical -- and the CPU oracle must equal both."""
import os

import numpy as np
import pytest

from cvgpuspeedup_b200 import _abi
from tests import gpu_util, util

pytestmark = pytest.mark.gpu

MUL, SUB, DIV = (0.3, 0.3, 0.3), (1.0, 4.0, 3.2), (3.2, 0.6, 11.8)


def _need(batch):
    if gpu_util.fkref_lib(batch) is None:
        pytest.skip(f"oracle/_ref/libfkref_{batch}.so not built (needs /root/reference at build time)")


def _ops(swap, mul=MUL, sub=SUB, div=DIV):
    return ([("reorder", (2, 1, 0))] if swap else []) + [("mul", mul), ("sub", sub), ("div", div)]


@pytest.mark.parametrize("swap", [0, 1])
@pytest.mark.parametrize("aspect", [_abi.IGNORE_AR, _abi.PRESERVE_AR, _abi.PRESERVE_AR_RN_EVEN, _abi.PRESERVE_AR_LEFT])
def test_reference_kernel_equals_ours_and_oracle(swap, aspect):
    _need(16)
    rng = np.random.default_rng(100 + aspect)
    img = util.make_image(rng, 640, 360, pitch=2048)
    rects = [(0, 0, 640, 360), (5, 7, 24, 48), (100, 3, 199, 33), (17, 150, 7, 5), (600, 0, 40, 360),
             (1, 1, 64, 128), (300, 100, 256, 200), (639, 359, 1, 1), (0, 0, 30, 120), (20, 20, 60, 120),
             (333, 111, 77, 99), (2, 300, 500, 60)]
    bg = (128.0, 3.5, 250.0)
    ref = gpu_util.run_fkref(img, rects, (64, 128), swap, MUL, SUB, DIV, aspect=aspect, bg=bg, batch=16, used=12)
    for variant in (1, 0):
        ours = gpu_util.run_cvgs(img, rects, (64, 128), _ops(swap), n_planes=16, used=12, variant=variant,
                                 aspect=aspect, background=bg)
        util.assert_bit_equal(ours, ref, f"ours(variant {variant}) vs reference kernel")
    orc = util.run_oracle(img, rects, (64, 128), _ops(swap), n_planes=16, used=12, aspect=aspect, background=bg)
    util.assert_bit_equal(orc, ref, "oracle vs reference kernel")


@pytest.mark.parametrize("src_type", [_abi.CVGS_16UC3, _abi.CVGS_16SC3])
@pytest.mark.parametrize("aspect", [_abi.IGNORE_AR, _abi.PRESERVE_AR])
def test_reference_kernel_16bit_sources(src_type, aspect):
    """ushort3 / short3 instantiations of the reference's Resize (its test matrix, test_batchresize_x_split3D.cu:427-432)."""
    _need(16)
    rng = np.random.default_rng(7 + src_type)
    img = rng.integers(0, 256, size=(240, 1936), dtype=np.uint8)  # 320 x 240 16-bit pixels, full value range
    rects = [(0, 0, 320, 240), (5, 7, 24, 48), (100, 3, 199, 33), (17, 150, 7, 5), (300, 0, 20, 240), (1, 1, 64, 128),
             (319, 239, 1, 1), (20, 20, 60, 120)]
    bg = (128.0, 3.5, 250.0)
    mul = (1 / 257.0,) * 3
    ref = gpu_util.run_fkref(img, rects, (64, 128), 1, mul, SUB, DIV, aspect=aspect, bg=bg, batch=16, used=8, src_type=src_type)
    ours = gpu_util.run_cvgs(img, rects, (64, 128), _ops(1, mul=mul), n_planes=16, used=8, aspect=aspect, background=bg,
                             src_type=src_type)
    util.assert_bit_equal(ours, ref, "16-bit source: ours vs reference kernel")
    orc = util.run_oracle(img, rects, (64, 128), _ops(1, mul=mul), n_planes=16, used=8, aspect=aspect, background=bg,
                          src_type=src_type)
    util.assert_bit_equal(orc, ref, "16-bit source: oracle vs reference kernel")


def test_reference_kernel_c2_fifty_crops():
    """BASELINE config 2 against the reference instantiated with BATCH=50."""
    _need(50)
    w = util.workload_c2()
    ref = gpu_util.run_fkref(w.image, w.rects, w.dsize, 1, MUL, SUB, DIV, batch=50)
    ours = gpu_util.run_cvgs(w.image, w.rects, w.dsize, w.ops)
    util.assert_bit_equal(ours, ref, "C2 vs reference kernel")


def test_reference_kernel_imagenet_chain():
    """BASELINE config 3 chain (scale 1/255, mean, std) on 128 crops -> 224x224, reference BATCH=128."""
    _need(128)
    w = util.workload_c3(n=128)
    ref = gpu_util.run_fkref(w.image, w.rects, w.dsize, 1, (1 / 255.0,) * 3, util._MEAN, util._STD, batch=128)
    ours = gpu_util.run_cvgs(w.image, w.rects, w.dsize, w.ops)
    util.assert_bit_equal(ours, ref, "C3 vs reference kernel")
