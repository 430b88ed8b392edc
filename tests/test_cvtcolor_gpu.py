"""Colour conversions that change the channel count, fused behind the resize (add / drop alpha,
-> gray; reference color_conversion.cuh:42-68,364-461, cvGS::cvtColor include/cvGPUSpeedup.cuh:151-161): bit-exact
against the reference's own kernel for the eight codes it supports, and against the oracle for chains, layouts,
batches and the 16-bit sources."""
import ctypes as C

import numpy as np
import pytest
import torch

from cvgpuspeedup_b200 import _abi
from tests import gpu_util, util

pytestmark = pytest.mark.gpu

MUL, SUB = (0.3, 1.7, 0.05, 0.25), (1.0, -4.0, 3.2, 0.5)


def _img(seed, nc, w=320, h=240):
    rng = np.random.default_rng(seed)
    return rng.integers(0, 256, size=(h, w * nc + 64), dtype=np.uint8)


@pytest.mark.parametrize("code", sorted(gpu_util.CVT_CODES))
def test_conversion_matches_reference_kernel_and_oracle(code):
    if gpu_util.fkref_lib(16) is None:
        pytest.skip("oracle/_ref not built")
    nc, cvt = gpu_util.CVT_CODES[code]
    src_type = _abi.CVGS_8UC3 if nc == 3 else _abi.CVGS_8UC4
    img = _img(900 + code, nc)
    nco = util.out_channels(src_type, cvt)
    ops = cvt + [("mul", MUL[:nco]), ("sub", SUB[:nco])]
    # (320, 240) is scale 1: the exact pixel values reach the conversion, and ~0.03 % of random (R, G, B) triples round
    # to different integers under the two FMUL orders nvcc picked for the gray codes (see CVGS_OP_GRAY)
    for dsize in [(64, 128), (400, 300), (33, 7), (320, 240)]:
        ref = gpu_util.run_fkref_cvt(code, img, 320, 240, dsize, MUL, SUB)
        ours = gpu_util.run_cvgs(img, [(0, 0, 320, 240)], dsize, ops, src_type=src_type)[0]
        orc = util.run_oracle(img, [(0, 0, 320, 240)], dsize, ops, src_type=src_type)[0]
        assert ours.shape == ref.shape == (nco, dsize[1], dsize[0])
        util.assert_bit_equal(ours, ref, f"code {code} dsize {dsize}: ours vs reference kernel")
        util.assert_bit_equal(orc, ref, f"code {code} dsize {dsize}: oracle vs reference kernel")


RECTS = [(0, 0, 320, 240), (5, 7, 24, 48), (100, 3, 199, 33), (17, 150, 7, 5), (300, 0, 20, 240), (319, 239, 1, 1)]
CHAINS = [
    [("mul", MUL[:3]), ("add_alpha", (255.0,)), ("sub", SUB)],                       # the FMA spans the conversion
    [("mul", MUL[:3]), ("reorder", (2, 1, 0)), ("add_alpha", (1.0,)), ("add", SUB), ("div", (3.2, 0.6, 11.8, 2.0))],
    [("mul", MUL[:3]), ("gray", (1,)), ("sub", (0.5,)), ("div", (0.25,))],
    [("sub", SUB[:3]), ("reorder", (1, 2, 0)), ("gray", (0,)), ("mul", (2.0,)), ("add", (1.0,))],
    [("add_alpha", (7.0,)), ("reorder", (3, 0, 1, 2)), ("mul", MUL), ("drop_alpha", ()), ("sub", SUB[:3])],
    [("add_alpha", (255.0,)), ("gray", ())],
]


@pytest.mark.parametrize("chain", range(len(CHAINS)))
@pytest.mark.parametrize("src_type", [_abi.CVGS_8UC3, _abi.CVGS_16UC3, _abi.CVGS_16SC3])
def test_chains_from_3_channel_sources(chain, src_type):
    ops = CHAINS[chain]
    img = _img(950 + chain, util.px_bytes_of(src_type))
    for kw in [dict(), dict(layout=_abi.OUT_NHWC, n_planes=8, used=6, background=(9.0, 1.0, 200.0)),
               dict(layout=_abi.OUT_CNHW, aspect=_abi.PRESERVE_AR, background=(3.0, 2.0, 1.0)),
               dict(fp_contract=_abi.FP_SEPARATE, interp_mode=_abi.INTERP_ROUND_U8)]:
        got = gpu_util.run_cvgs(img, RECTS, (61, 45), ops, src_type=src_type, **kw)
        want = util.run_oracle(img, RECTS, (61, 45), ops, src_type=src_type, **kw)
        util.assert_bit_equal(got, want, f"chain {chain} src {src_type} {kw}")


CHAINS4 = [
    [("drop_alpha", ()), ("mul", MUL[:3]), ("sub", SUB[:3])],
    [("mul", MUL), ("reorder", (2, 1, 0, 3)), ("drop_alpha", ()), ("sub", SUB[:3])],
    [("reorder", (3, 2, 1, 0)), ("gray", ()), ("mul", (0.5,))],
    [("drop_alpha", ()), ("add_alpha", (255.0,)), ("mul", MUL)],
]


@pytest.mark.parametrize("chain", range(len(CHAINS4)))
@pytest.mark.parametrize("src_type", [_abi.CVGS_8UC4, _abi.CVGS_16UC4])
def test_chains_from_4_channel_sources(chain, src_type):
    ops = CHAINS4[chain]
    img = _img(970 + chain, util.px_bytes_of(src_type))
    for kw in [dict(), dict(layout=_abi.OUT_NHWC, n_planes=7, used=5, background=(9.0, 1.0, 200.0, 4.0)),
               dict(layout=_abi.OUT_CNHW, fp_contract=_abi.FP_SEPARATE)]:
        got = gpu_util.run_cvgs(img, RECTS, (40, 52), ops, src_type=src_type, **kw)
        want = util.run_oracle(img, RECTS, (40, 52), ops, src_type=src_type, **kw)
        util.assert_bit_equal(got, want, f"chain {chain} src {src_type} {kw}")


def test_large_batch_goes_through_the_descriptor_ring():
    img = _img(990, 3, 640, 480)
    rng = np.random.default_rng(3)
    rects = [(int(rng.integers(0, 500)), int(rng.integers(0, 300)), int(rng.integers(8, 140)), int(rng.integers(8, 180)))
             for _ in range(300)]
    ops = [("reorder", (2, 1, 0)), ("gray", ()), ("mul", (1 / 255.0,))]
    got = gpu_util.run_cvgs(img, rects, (32, 32), ops)
    want = util.run_oracle(img, rects, (32, 32), ops)
    assert got.shape == (300, 1, 32, 32)
    util.assert_bit_equal(got, want, "300 crops -> gray")


def test_bad_chains_are_rejected():
    lib = _abi.load()
    t = torch.zeros(64 * 64 * 4, dtype=torch.uint8, device="cuda")
    out = torch.zeros(4 * 8 * 8, device="cuda")
    crops = (_abi.Crop * 1)()
    crops[0].data, crops[0].width, crops[0].height, crops[0].pitch = t.data_ptr(), 64, 64, 256

    def rc(ops, **kw):
        p = util.make_pipeline((8, 8), ops, out_ptr=out.data_ptr(), **kw)
        return lib.cvgs_b200_preproc_launch(crops, 1, 1, C.byref(p), None)

    assert rc([("drop_alpha", ())]) != 0 and b"4-channel" in lib.cvgs_b200_last_error()
    assert rc([("add_alpha", (1.0,))], src_type=_abi.CVGS_8UC4) != 0
    assert rc([("gray", ()), ("gray", ())]) != 0
    assert rc([("gray", ())], dst_type=_abi.CVGS_32FC3) != 0 and b"dst_type" in lib.cvgs_b200_last_error()
    assert rc([("gray", ())], dst_type=_abi.CVGS_32FC1) == 0
    assert rc([("gray", ())], layout=_abi.OUT_PLANES) != 0
    assert rc([("gray", ())], layout=_abi.OUT_NHWC, dst_type=_abi.CVGS_8UC3) != 0
    assert rc([("add_alpha", (1.0,)), ("reorder", (0, 1, 2, 4))]) != 0
    torch.cuda.synchronize()


def test_python_api_codes():
    import cvgpuspeedup_b200 as cvgs
    rng = np.random.default_rng(77)
    img = util.make_image(rng, 160, 120, 512)
    t = gpu_util.device_image(img)
    mat = cvgs.GpuMat(t.data_ptr(), 160, 120, 512, owner=t)
    out4 = torch.full((1, 4, 60, 80), float("nan"), device="cuda")
    cvgs.executeOperations(None, cvgs.resize([mat], (80, 60)), cvgs.cvtColor(cvgs.api.COLOR_BGR2RGBA),
                           cvgs.multiply((0.5, 0.25, 2.0, 1.0)), cvgs.split(out4))
    out1 = torch.full((1, 1, 60, 80), float("nan"), device="cuda")
    cvgs.executeOperations(None, cvgs.resize([mat], (80, 60)), cvgs.cvtColor(cvgs.api.COLOR_RGB2GRAY), cvgs.split(out1))
    torch.cuda.synchronize()
    want4 = util.run_oracle(img, [(0, 0, 160, 120)], (80, 60),
                            [("reorder", (2, 1, 0)), ("add_alpha", (255.0,)), ("mul", (0.5, 0.25, 2.0, 1.0))])
    want1 = util.run_oracle(img, [(0, 0, 160, 120)], (80, 60), [("gray", (1,))])
    util.assert_bit_equal(out4.cpu().numpy(), want4, "BGR2RGBA")
    util.assert_bit_equal(out1.cpu().numpy(), want1, "RGB2GRAY")


def test_unused_planes_take_the_converted_default():
    """Planes z >= used: the chain -- conversion included -- applied to the default value (BatchRead default + chain)."""
    img = _img(995, 3)
    ops = [("mul", (2.0, 3.0, 4.0)), ("reorder", (2, 1, 0)), ("add_alpha", (255.0,)), ("sub", (1.0, 1.0, 1.0, 5.0))]
    got = gpu_util.run_cvgs(img, [(0, 0, 320, 240)], (16, 12), ops, n_planes=3, used=1, background=(10.0, 20.0, 30.0))
    want = util.run_oracle(img, [(0, 0, 320, 240)], (16, 12), ops, n_planes=3, used=1, background=(10.0, 20.0, 30.0))
    util.assert_bit_equal(got, want, "defaults through a conversion")
    assert got[2, :, 0, 0].tolist() == [119.0, 59.0, 19.0, 250.0]
    got0 = gpu_util.run_cvgs(img, [], (16, 12), [("gray", (0,))], n_planes=2, used=0, background=(100.0, 100.0, 100.0))
    assert got0.shape == (2, 1, 12, 16) and (got0 == 100.0).all()


@pytest.mark.parametrize("n,parents", [(20, True), (100, True), (300, True), (20, False), (300, False)])
@pytest.mark.parametrize("ops", [
    [("gray", (1,))],                                                            # RGB2GRAY
    [("reorder", (2, 1, 0)), ("gray", (0,)), ("mul", (1 / 255.0,))],             # BGR2GRAY + scale
    [("gray", (1,)), ("mul", (0.5,)), ("sub", (3.0,)), ("div", (7.0,))],         # RGB2GRAY + a whole chain on the one channel
])
def test_tma_staged_kernel_takes_gray_chains(n, parents, ops):
    """cvtColor<*2GRAY> first in the chain, one float plane out, crops with their parent frame named: the TMA-staged
    kernel's CH_GRAY instantiation (forced: variant 2 fails instead of falling back), bit-equal to the oracle in both
    floating-point contracts."""
    import ctypes as C
    lib = _abi.load()
    w = util.workload_c2(seed=50 + n, n=n, frame=(960, 540), pitch=2880)
    d_img = torch.from_numpy(w.image).cuda()
    for kw in ({}, dict(fp_contract=_abi.FP_SEPARATE)):
        out = torch.full((n, 1, 128, 64), float("nan"), device="cuda")
        p = util.make_pipeline((64, 128), ops, out_ptr=out.data_ptr(), **kw)
        crops = util.host_crops(w.image, w.rects, base_ptr=d_img.data_ptr())
        par = util.host_parents(w.image, w.width, w.height, n, base_ptr=d_img.data_ptr()) if parents else None
        prev = lib.cvgs_b200_set_kernel_variant(2)
        try:
            _abi.check(lib.cvgs_b200_preproc_launch_ex(crops, par, n, n, C.byref(p), None))
        finally:
            lib.cvgs_b200_set_kernel_variant(prev)
        torch.cuda.synchronize()
        util.assert_bit_equal(out.cpu().numpy(), util.run_oracle(w.image, w.rects, (64, 128), ops, **kw), f"gray chain {ops} {kw}")


@pytest.mark.parametrize("n,parents", [(20, True), (100, True), (300, True), (300, False), (20, False), (64, False)])
@pytest.mark.parametrize("ops", [
    [("add_alpha", (255.0,))],                                                                          # RGB2RGBA
    [("reorder", (2, 1, 0)), ("add_alpha", (255.0,)), ("mul", (1 / 255.0,) * 4), ("sub", (0.485, 0.456, 0.406, 0.5)),
     ("div", (0.229, 0.224, 0.225, 0.25))],                                                             # BGR2RGBA + normalisation
    [("mul", MUL[:3]), ("add_alpha", (1.0,)), ("sub", SUB)],                                            # the FMA spans the conversion
    [("add_alpha", (7.0,)), ("reorder", (3, 0, 1, 2)), ("mul", MUL), ("add", SUB), ("mul", (2.0, 0.5, 1.5, 3.0))],  # alpha first, generic chain
])
def test_tma_staged_kernel_takes_alpha_chains(n, parents, ops):
    """cvtColor<*2*A> in the chain of a CV_8UC3 source, four float planes out: the TMA-staged kernel's CH_*_ALPHA
    instantiations (forced: variant 2 fails instead of falling back) -- the chain on the three source channels, the alpha plane
    filled with the value the host computed by running the constant through the same ops.  Bit-equal to the oracle in both
    floating-point contracts, NCHW and CNHW."""
    import ctypes as C
    lib = _abi.load()
    w = util.workload_c2(seed=60 + n, n=n, frame=(960, 540), pitch=2880)
    d_img = torch.from_numpy(w.image).cuda()
    for kw in ({}, dict(fp_contract=_abi.FP_SEPARATE), dict(layout=_abi.OUT_CNHW)):
        shape = (4, n, 128, 64) if kw.get("layout") == _abi.OUT_CNHW else (n, 4, 128, 64)
        out = torch.full(shape, float("nan"), device="cuda")
        p = util.make_pipeline((64, 128), ops, out_ptr=out.data_ptr(), **kw)
        crops = util.host_crops(w.image, w.rects, base_ptr=d_img.data_ptr())
        par = util.host_parents(w.image, w.width, w.height, n, base_ptr=d_img.data_ptr()) if parents else None
        prev = lib.cvgs_b200_set_kernel_variant(2)
        try:
            _abi.check(lib.cvgs_b200_preproc_launch_ex(crops, par, n, n, C.byref(p), None))
        finally:
            lib.cvgs_b200_set_kernel_variant(prev)
        torch.cuda.synchronize()
        util.assert_bit_equal(out.cpu().numpy(), util.run_oracle(w.image, w.rects, (64, 128), ops, **kw), f"alpha chain {ops} {kw}")
