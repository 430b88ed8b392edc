"""CPU checks of the colour conversions that change the channel count: the oracle against golden vectors from the
reference's own kernel (tests/golden/cvt_code*.npz, made by tests/golden/make_golden_cvt.py on a B200) and known
answers."""
import glob
import os

import numpy as np
import pytest

from cvgpuspeedup_b200 import _abi
from tests import util

GOLDEN = sorted(glob.glob(os.path.join(util.ROOT, "tests", "golden", "cvt_code*.npz")))
# cv:: / fk:: ColorConversionCodes -> ops of the C-ABI (same table as tests/gpu_util.py, which needs torch)
CVT_OPS = {0: [("add_alpha", (255.0,))], 1: [("drop_alpha", ())], 2: [("reorder", (2, 1, 0)), ("add_alpha", (255.0,))],
           3: [("reorder", (2, 1, 0, 3)), ("drop_alpha", ())], 6: [("reorder", (2, 1, 0)), ("gray", (0,))], 7: [("gray", (1,))],
           10: [("reorder", (2, 1, 0, 3)), ("gray", (0,))], 11: [("gray", (1,))]}


@pytest.mark.parametrize("path", GOLDEN or [None])
def test_oracle_matches_reference_kernel_golden(path):
    if path is None:
        pytest.skip("no colour-conversion golden vectors committed")
    g = np.load(path)
    code, nc = int(g["code"]), int(g["channels"])
    src_type = _abi.CVGS_8UC3 if nc == 3 else _abi.CVGS_8UC4
    img, w, h = np.ascontiguousarray(g["image"]), int(g["width"]), int(g["height"])
    nco = util.out_channels(src_type, CVT_OPS[code])
    ops = CVT_OPS[code] + [("mul", tuple(float(v) for v in g["mul"][:nco])), ("sub", tuple(float(v) for v in g["sub"][:nco]))]
    for key in [k for k in g.files if k.startswith("out_")]:
        dw, dh = (int(v) for v in key[4:].split("x"))
        got = util.run_oracle(img, [(0, 0, w, h)], (dw, dh), ops, src_type=src_type)[0]
        util.assert_bit_equal(got, g[key], f"{os.path.basename(path)} {key}")


def test_known_answers():
    rng = np.random.default_rng(5)
    img = util.make_image(rng, 20, 10, 64)
    px = img[:, :60].reshape(10, 20, 3).astype(np.float32)
    rgba = util.run_oracle(img, [(0, 0, 20, 10)], (20, 10), [("add_alpha", (255.0,))], layout=_abi.OUT_NHWC)[0]
    assert rgba.shape == (10, 20, 4) and np.array_equal(rgba[..., :3], px) and (rgba[..., 3] == 255).all()
    bgra = util.run_oracle(img, [(0, 0, 20, 10)], (20, 10), [("reorder", (2, 1, 0)), ("add_alpha", (255.0,))], layout=_abi.OUT_NHWC)[0]
    assert np.array_equal(bgra[..., :3], px[..., ::-1])
    gray = util.run_oracle(img, [(0, 0, 20, 10)], (20, 10), [("gray", (1,))])[0, 0]
    lum = px[..., 0].astype(np.float64) * 0.299 + px[..., 1] * 0.587 + px[..., 2] * 0.114
    assert np.array_equal(gray, np.rint(gray))            # the reference rounds the luminance (see CVGS_OP_GRAY)
    assert np.abs(gray - lum).max() <= 0.5 + 1e-3
    # a 4-channel source: drop the alpha, then the rest of the chain sees three channels
    img4 = rng.integers(0, 256, size=(10, 20 * 4), dtype=np.uint8)
    rgb = util.run_oracle(img4, [(0, 0, 20, 10)], (20, 10), [("drop_alpha", ()), ("mul", (2.0, 3.0, 4.0))],
                          src_type=_abi.CVGS_8UC4, layout=_abi.OUT_NHWC)[0]
    assert np.array_equal(rgb, img4.reshape(10, 20, 4)[..., :3].astype(np.float32) * np.array([2, 3, 4], dtype=np.float32))


def test_mul_add_contracts_across_the_alpha_conversion():
    """(x*m) then AddOpaqueAlpha then -s: still one FMA per colour channel in the reference's inlined chain."""
    img = np.full((4, 16), 5, dtype=np.uint8)
    m, s = np.float32(0.1), np.float32(0.485)
    out = util.run_oracle(img, [(0, 0, 4, 4)], (4, 4), [("mul", (0.1, 0.1, 0.1)), ("add_alpha", (255.0,)), ("sub", (0.485,) * 4)])[0]
    fused = np.float32(np.float64(np.float32(5)) * np.float64(m) - np.float64(s))
    assert out[0, 0, 0] == fused and out[0, 0, 0] != np.float32(np.float32(5) * m) - s
    assert out[3, 0, 0] == np.float32(255) - s
    sep = util.run_oracle(img, [(0, 0, 4, 4)], (4, 4), [("mul", (0.1, 0.1, 0.1)), ("add_alpha", (255.0,)), ("sub", (0.485,) * 4)],
                          fp_contract=_abi.FP_SEPARATE)[0]
    assert sep[0, 0, 0] == np.float32(np.float32(5) * m) - s
