"""CPU checks of the warp path's oracle (oracle/oracle.c oracle_warp) and host logic: the reference-kernel golden
vectors under tests/golden/warp_*.npz (tests/golden/make_golden_warp.py), identities, matrix inversion."""
import ctypes as C
import glob
import os

import numpy as np
import pytest

import cvgpuspeedup_b200 as cvgs
from cvgpuspeedup_b200 import _abi
from tests import util

GOLDEN = sorted(glob.glob(os.path.join(util.ROOT, "tests", "golden", "warp_*.npz")))


def _oracle(img, w, h, inv, warp_type, dsize, ops, u8=None, n_planes=1, used=1, background=(0, 0, 0)):
    crops = (_abi.Crop * 1)()
    warps = (_abi.Warp * 1)()
    crops[0].data, crops[0].width, crops[0].height, crops[0].pitch = img.ctypes.data, w, h, img.shape[1]
    warps[0].type = int(warp_type)
    for k in range(9):
        warps[0].m[k] = float(inv[k])
    if u8 is None:
        out = np.full((n_planes, 3, dsize[1], dsize[0]), np.nan, dtype=np.float32)
        p = util.make_pipeline(dsize, ops, out_ptr=out.ctypes.data, background=background)
    else:
        out = np.full((n_planes, dsize[1], dsize[0], 3), 99, dtype=np.uint8)
        p = util.make_pipeline(dsize, ops, out_ptr=out.ctypes.data, layout=_abi.OUT_NHWC, dst_type=_abi.CVGS_8UC3,
                               u8_cast=u8, background=background)
    assert util.oracle_lib().oracle_warp(crops, warps, n_planes, used, C.byref(p), 1) == 0
    return out


@pytest.mark.parametrize("path", GOLDEN or [None])
def test_oracle_warp_matches_reference_kernel_golden(path):
    if path is None:
        pytest.skip("no warp golden vectors committed")
    g = np.load(path)
    img, w, h = np.ascontiguousarray(g["image"]), int(g["width"]), int(g["height"])
    dsize = (g["out_f32"].shape[2], g["out_f32"].shape[1])
    f = _oracle(img, w, h, g["inverse"], int(g["warp_type"]), dsize, [("mul", tuple(float(v) for v in g["mul"]))])[0]
    util.assert_bit_equal(f, g["out_f32"], os.path.basename(path))
    u = _oracle(img, w, h, g["inverse"], int(g["warp_type"]), dsize, [], u8=1)[0]
    assert np.array_equal(u, g["out_u8"])


def test_identity_and_integer_shift():
    rng = np.random.default_rng(1)
    w, h = 37, 23
    img = util.make_image(rng, w, h, 128)
    px = img[:, :3 * w].reshape(h, w, 3)
    ident = cvgs.api.invert_warp_matrix(np.array([[1, 0, 0], [0, 1, 0]]), cvgs.WARP_AFFINE)
    assert np.array_equal(_oracle(img, w, h, ident, cvgs.WARP_AFFINE, (w, h), [], u8=1)[0], px)
    shift = cvgs.api.invert_warp_matrix(np.array([[1, 0, 4], [0, 1, 6]]), cvgs.WARP_AFFINE)
    got = _oracle(img, w, h, shift, cvgs.WARP_AFFINE, (w, h), [], u8=1)[0]
    assert np.array_equal(got[6:, 4:], px[:h - 6, :w - 4])
    assert not got[:6].any() and not got[:, :4].any()  # outside the source: zeros (warping.cuh:80-82)
    # a perspective matrix with an affine last row gives the affine result
    p = np.array([[1, 0, 4], [0, 1, 6], [0, 0, 1]], dtype=np.float64)
    pin = cvgs.api.invert_warp_matrix(p, cvgs.WARP_PERSPECTIVE)
    assert np.array_equal(_oracle(img, w, h, pin, cvgs.WARP_PERSPECTIVE, (w, h), [], u8=1)[0], got)


def test_unused_planes_take_the_default_through_the_chain():
    rng = np.random.default_rng(2)
    img = util.make_image(rng, 16, 16, 64)
    ident = cvgs.api.invert_warp_matrix(np.array([[1, 0, 0], [0, 1, 0]]), cvgs.WARP_AFFINE)
    out = _oracle(img, 16, 16, ident, cvgs.WARP_AFFINE, (8, 8), [("mul", (2.0, 2.0, 2.0))], n_planes=2, used=1,
                  background=(1.0, 2.0, 3.0))
    assert np.array_equal(out[1, :, 0, 0], np.array([2.0, 4.0, 6.0], dtype=np.float32))


def test_invert_warp_matrix():
    m = np.array([[0.8, 0.3, -20.5], [-0.25, 1.1, 33.25]])
    inv = cvgs.api.invert_warp_matrix(m, cvgs.WARP_AFFINE).astype(np.float64)
    full = np.vstack([m, [0, 0, 1]])
    assert np.allclose(np.vstack([inv[:6].reshape(2, 3), [0, 0, 1]]) @ full, np.eye(3), atol=1e-5)
    assert not cvgs.api.invert_warp_matrix(np.zeros((2, 3)), cvgs.WARP_AFFINE).any()  # singular: zeros, like OpenCV
    h = np.array([[1.1, 0.1, 3], [0.05, 0.9, -2], [1e-4, 2e-4, 1]])
    hin = cvgs.api.invert_warp_matrix(h, cvgs.WARP_PERSPECTIVE).astype(np.float64).reshape(3, 3)
    assert np.allclose(hin @ h, np.eye(3), atol=1e-5)
    with pytest.raises(cvgs.CvgsError):
        cvgs.api.invert_warp_matrix(np.eye(3), cvgs.WARP_AFFINE)


def test_oracle_warp_rejects_bad_arguments():
    lib = util.oracle_lib()
    p = util.make_pipeline((4, 4), [], out_ptr=0)
    assert lib.oracle_warp(None, None, 1, 1, C.byref(p), 1) != 0
