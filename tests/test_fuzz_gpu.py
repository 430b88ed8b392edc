"""Randomised geometry against the oracle: crop sizes / positions / pitches / base alignments, destination sizes,
aspect modes, layouts, used < planes, op chains, both ways of naming sources (per-crop tensor maps and per-image
maps), the automatic kernel choice and the forced TMA kernel.  Seeds are fixed: failures reproduce."""
import numpy as np
import pytest
import torch

from cvgpuspeedup_b200 import _abi
from tests import gpu_util, util

pytestmark = pytest.mark.gpu

OPS_POOL = [
    [],
    [("mul", (0.5, 0.25, 2.0))],
    [("reorder", (2, 1, 0)), ("mul", (0.3, 0.3, 0.3)), ("sub", (1.0, 4.0, 3.2)), ("div", (3.2, 0.6, 11.8))],
    [("div", (255.0, 255.0, 255.0)), ("sub", (0.485, 0.456, 0.406)), ("div", (0.229, 0.224, 0.225))],
    [("mul", (1 / 255.0,) * 3), ("sub", (0.485, 0.456, 0.406)), ("div", (-0.229, 0.224, 7.0))],
    [("add", (1.5, -2.5, 3.5)), ("reorder", (1, 2, 0)), ("mul", (2.0, 3.0, -4.0)), ("add", (0.1, 0.2, 0.3))],
    [("sub", (127.5,) * 3), ("div", (127.5,) * 3)],
]


@pytest.mark.parametrize("seed", range(24))
def test_random_geometry(seed):
    rng = np.random.default_rng(1000 + seed)
    fw, fh = int(rng.integers(40, 700)), int(rng.integers(30, 500))
    pitch = (3 * fw + 15) // 16 * 16 + 16 * int(rng.integers(0, 5))
    shift = int(rng.integers(0, 16)) if seed % 3 == 0 else 0          # unaligned image base
    back = rng.integers(0, 256, size=fh * pitch + 64, dtype=np.uint8)
    img = back[shift:shift + fh * pitch].reshape(fh, pitch)
    d_back = torch.from_numpy(back).cuda()
    d_img = d_back[shift:shift + fh * pitch].view(fh, pitch)
    n = int(rng.integers(1, 70 if seed % 4 else 140))
    rects = []
    for _ in range(n):
        w = int(rng.integers(1, fw + 1)) if rng.random() < 0.7 else int(rng.integers(1, min(fw, 12) + 1))
        h = int(rng.integers(1, fh + 1)) if rng.random() < 0.7 else int(rng.integers(1, min(fh, 12) + 1))
        rects.append((int(rng.integers(0, fw - w + 1)), int(rng.integers(0, fh - h + 1)), w, h))
    dsize = (int(rng.integers(1, 300)), int(rng.integers(1, 200)))
    ops = OPS_POOL[int(rng.integers(0, len(OPS_POOL)))]
    aspect = int(rng.integers(0, 4))
    layout = int(rng.integers(0, 3))
    n_planes = n + int(rng.integers(0, 3))
    kw = dict(aspect=aspect, background=(float(rng.integers(0, 255)), 3.5, 200.0), layout=layout, n_planes=n_planes, used=n)
    if rng.random() < 0.3:
        kw.update(fp_contract=_abi.FP_SEPARATE)
    if rng.random() < 0.3:
        kw.update(interp_mode=_abi.INTERP_ROUND_U8)
    want = util.run_oracle(img, rects, dsize, ops, **kw)
    for variant, parents in [(0, None), (0, (fw, fh)), (1, None)]:
        got = gpu_util.run_cvgs(img, rects, dsize, ops, variant=variant, d_image=d_img, parents=parents, **kw)
        util.assert_bit_equal(got, want, f"seed {seed} variant {variant} parents {parents} frame {fw}x{fh} pitch {pitch} "
                                         f"shift {shift} n {n} dsize {dsize} aspect {aspect} layout {layout}")
