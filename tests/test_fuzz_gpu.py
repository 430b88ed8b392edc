"""Randomised geometry against the oracle: crop sizes / positions / pitches / base alignments, destination sizes,
aspect modes, layouts, used < planes, op chains, both ways of naming sources (per-crop tensor maps and per-image
maps), the automatic kernel choice and the forced TMA kernel.  Seeds are fixed: failures reproduce."""
import numpy as np
import pytest
import torch

from cvgpuspeedup_b200 import _abi
from tests import gpu_util, util

pytestmark = pytest.mark.gpu
# CVGS_FUZZ_EXTRA=n adds n more seeds to every fuzz test (one-off soak runs; the default suite stays short)
import os
_EXTRA = int(os.environ.get("CVGS_FUZZ_EXTRA", "0"))

OPS_POOL = [
    [],
    [("mul", (0.5, 0.25, 2.0))],
    [("reorder", (2, 1, 0)), ("mul", (0.3, 0.3, 0.3)), ("sub", (1.0, 4.0, 3.2)), ("div", (3.2, 0.6, 11.8))],
    [("div", (255.0, 255.0, 255.0)), ("sub", (0.485, 0.456, 0.406)), ("div", (0.229, 0.224, 0.225))],
    [("mul", (1 / 255.0,) * 3), ("sub", (0.485, 0.456, 0.406)), ("div", (-0.229, 0.224, 7.0))],
    [("add", (1.5, -2.5, 3.5)), ("reorder", (1, 2, 0)), ("mul", (2.0, 3.0, -4.0)), ("add", (0.1, 0.2, 0.3))],
    [("sub", (127.5,) * 3), ("div", (127.5,) * 3)],
]


@pytest.mark.parametrize("seed", range(24 + _EXTRA))
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


def _random_chain(rng, nc):
    """A random chain over the per-channel ops, reorders and the channel-count changing conversions; returns
    (ops, channels of the result)."""
    ops = []
    for _ in range(int(rng.integers(0, 6))):
        k = int(rng.integers(0, 8))
        vals = tuple(float(v) for v in rng.uniform(0.25, 4.0, size=nc) * rng.choice([-1.0, 1.0], size=nc))
        if k == 0:
            ops.append(("mul", vals))
        elif k == 1:
            ops.append(("sub", vals))
        elif k == 2:
            ops.append(("add", vals))
        elif k == 3:
            ops.append(("div", vals))
        elif k == 4 and nc >= 3:
            ops.append(("reorder", tuple(int(v) for v in rng.permutation(nc))))
        elif k == 5 and nc == 3 and not any(o[0] == "add_alpha" for o in ops):
            ops.append(("add_alpha", (float(rng.choice([255.0, 1.0, 0.5])),)))
            nc = 4
        elif k == 6 and nc == 4:
            ops.append(("drop_alpha", ()))
            nc = 3
        elif k == 7 and nc >= 3:
            ops.append(("gray", (int(rng.integers(0, 2)),)))
            nc = 1
    return ops, nc


SRC_TYPES = [_abi.CVGS_8UC3, _abi.CVGS_16UC3, _abi.CVGS_16SC3, _abi.CVGS_8UC4, _abi.CVGS_16UC4, _abi.CVGS_16SC4]


@pytest.mark.parametrize("seed", range(36 + _EXTRA))
def test_random_source_types_chains_and_outputs(seed):
    """The forms outside the TMA kernel: every source depth / channel count, chains with the colour conversions,
    8-bit output, packed float output with padded rows -- against the oracle."""
    import ctypes as C
    rng = np.random.default_rng(5000 + seed)
    src_type = SRC_TYPES[seed % len(SRC_TYPES)]
    px, nc = util.px_bytes_of(src_type), util.channels_of(src_type)
    fw, fh = int(rng.integers(20, 300)), int(rng.integers(20, 200))
    pitch = (px * fw + 15) // 16 * 16 + 16 * int(rng.integers(0, 3))
    img = rng.integers(0, 256, size=(fh, pitch), dtype=np.uint8)
    d_img = torch.from_numpy(img).cuda()
    n = int(rng.integers(1, 80))
    rects = []
    for _ in range(n):
        w, h = int(rng.integers(1, fw + 1)), int(rng.integers(1, fh + 1))
        rects.append((int(rng.integers(0, fw - w + 1)), int(rng.integers(0, fh - h + 1)), w, h))
    W, H = int(rng.integers(1, 150)), int(rng.integers(1, 100))
    ops, nco = _random_chain(rng, nc)
    n_planes = n + int(rng.integers(0, 3))
    kw = dict(aspect=int(rng.integers(0, 4)), background=tuple(float(v) for v in rng.uniform(-5, 260, size=4))[:nc],
              src_type=src_type)
    if rng.random() < 0.3:
        kw.update(fp_contract=_abi.FP_SEPARATE)
    if rng.random() < 0.3:
        kw.update(interp_mode=_abi.INTERP_ROUND_U8)
    form = int(rng.integers(0, 4))
    lib = _abi.load()
    crops_h = util.host_crops(img, rects, px_bytes=px)
    crops_d = util.host_crops(img, rects, base_ptr=d_img.data_ptr(), px_bytes=px)
    what = f"seed {seed} src {src_type} ops {ops} form {form} dsize {(W, H)} n {n}/{n_planes} {kw}"
    if form == 0 and nco >= 3:  # 8-bit packed output with its own pitch
        rp = nco * W + int(rng.integers(0, 9))
        kw.update(layout=_abi.OUT_NHWC, dst_type=_abi.CVGS_8UC3 if nco == 3 else _abi.CVGS_8UC4, row_pitch=rp,
                  u8_cast=0)
        want = np.full((n_planes, H, rp), 7, dtype=np.uint8)
        got = torch.full((n_planes, H, rp), 7, dtype=torch.uint8, device="cuda")
    elif form == 1:  # packed float output with padded rows
        rs = nco * W + int(rng.integers(0, 6))
        kw.update(layout=_abi.OUT_NHWC, row_pitch=4 * rs)
        want = np.full((n_planes, H, rs), -3.0, dtype=np.float32)
        got = torch.full((n_planes, H, rs), -3.0, dtype=torch.float32, device="cuda")
    else:  # tensor layouts
        layout = int(rng.integers(0, 3))
        kw.update(layout=layout)
        shape = util.out_shape(n_planes, (W, H), layout, 0, nco)
        want = np.full(shape, np.nan, dtype=np.float32)
        got = torch.full(shape, float("nan"), dtype=torch.float32, device="cuda")
    p = util.make_pipeline((W, H), ops, out_ptr=want.ctypes.data, **kw)
    assert util.oracle_lib().oracle_preproc(crops_h, n_planes, n, C.byref(p), 0) == 0, what
    p = util.make_pipeline((W, H), ops, out_ptr=got.data_ptr(), **kw)
    _abi.check(lib.cvgs_b200_preproc_launch(crops_d, n_planes, n, C.byref(p), None))
    torch.cuda.synchronize()
    if want.dtype == np.uint8:
        assert np.array_equal(got.cpu().numpy(), want), what
    else:
        util.assert_bit_equal(got.cpu().numpy(), want, what)


@pytest.mark.parametrize("seed", range(36 + _EXTRA))
def test_random_common_geometry_every_source_type(seed):
    """The geometry the TMA-staged instantiations of the wider pixel types are built for (IGNORE_AR, every plane used,
    planar float tensors or -- CV_8UC3 -- packed 8-bit pixels): every source type, chains with and without the alpha / gray
    conversions, image bases at random aligned offsets, with and without parent images, small and large batches.  The
    automatic kernel choice and the direct-gather kernel against the oracle."""
    rng = np.random.default_rng(9000 + seed)
    src_type = SRC_TYPES[seed % len(SRC_TYPES)]
    px, nc = util.px_bytes_of(src_type), util.channels_of(src_type)
    fw, fh = int(rng.integers(20, 500)), int(rng.integers(20, 300))
    pitch = (px * fw + 15) // 16 * 16 + 16 * int(rng.integers(0, 3))
    shift = 8 * int(rng.integers(0, 2)) if seed % 2 else 0
    back = rng.integers(0, 256, size=fh * pitch + 64, dtype=np.uint8)
    img = back[shift:shift + fh * pitch].reshape(fh, pitch)
    d_back = torch.from_numpy(back).cuda()
    d_img = d_back[shift:shift + fh * pitch].view(fh, pitch)
    n = int(rng.integers(1, 60)) if seed % 3 else int(rng.integers(65, 300))
    rects = []
    for _ in range(n):
        w, h = int(rng.integers(1, fw + 1)), int(rng.integers(1, fh + 1))
        rects.append((int(rng.integers(0, fw - w + 1)), int(rng.integers(0, fh - h + 1)), w, h))
    dsize = (int(rng.integers(1, 260)), int(rng.integers(1, 160)))
    if nc == 3 and seed % 4 == 1:   # alpha in the chain
        tail = [[], [("mul", (0.5, 0.25, 2.0, 1.5))], [("mul", (1 / 255.0,) * 4), ("sub", (0.4, 0.5, 0.6, 0.7)), ("div", (0.2, 0.3, 0.4, 0.5))],
                [("add", (1.0, 2.0, 3.0, 4.0)), ("mul", (2.0, 3.0, 4.0, 5.0)), ("sub", (1.0, 1.0, 1.0, 1.0))]][int(rng.integers(0, 4))]
        ops = [("reorder", tuple(int(v) for v in rng.permutation(3))), ("add_alpha", (float(rng.choice([255.0, 1.0])),))] + tail
    elif nc == 3 and seed % 4 == 3:  # gray first
        ops = [("gray", (int(rng.integers(0, 2)),))] + [[], [("mul", (1 / 255.0,))], [("sub", (3.0,)), ("div", (7.0,))]][int(rng.integers(0, 3))]
    else:
        ops = [(k, tuple(v[:nc]) if k != "reorder" else tuple(int(x) for x in rng.permutation(nc)))
               for (k, v) in [[], [("mul", (0.5, 0.25, 2.0, 1.5))], [("mul", (0.3,) * 4), ("sub", (1.0, 4.0, 3.2, 0.5)), ("div", (3.2, 0.6, 11.8, 2.0))],
                              [("reorder", ()), ("add", (1.5, -2.5, 3.5, 0.5)), ("mul", (2.0, 3.0, -4.0, 1.0)), ("add", (0.1, 0.2, 0.3, 0.4))],
                              [("sub", (127.5,) * 4), ("div", (127.5,) * 4)]][int(rng.integers(0, 5))]]
    kw = dict(src_type=src_type, layout=int(rng.choice([_abi.OUT_NCHW, _abi.OUT_CNHW])))
    if rng.random() < 0.25:
        kw.update(fp_contract=_abi.FP_SEPARATE)
    if rng.random() < 0.2:
        kw.update(interp_mode=_abi.INTERP_ROUND_U8)
    want = util.run_oracle(img, rects, dsize, ops, **kw)
    parents = (fw, fh) if src_type == _abi.CVGS_8UC3 or seed % 2 == 0 else None
    for variant in (0, 1):
        got = gpu_util.run_cvgs(img, rects, dsize, ops, variant=variant, d_image=d_img, parents=parents if variant == 0 else None, **kw)
        util.assert_bit_equal(got, want, f"seed {seed} src {src_type} variant {variant} frame {fw}x{fh} pitch {pitch} shift {shift} "
                                         f"n {n} dsize {dsize} ops {ops} {kw}")
