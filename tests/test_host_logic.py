"""Host-side logic of the launch path, without a GPU: the TMA kernel's plan (band width, staging slots, grid, the
cost-weighted partition of work items over warps) and the bookkeeping that decides whether consecutive launches may
overlap.  Both are exported as diagnostics of the C-ABI so that the real code is what runs here."""
import ctypes as C

import numpy as np
import pytest

from cvgpuspeedup_b200 import _abi
from tests import util


def _plan(rects, dsize, n_planes=None, pitch=6144, image_mode=1, items_per_warp=1, sms=148, aspect=_abi.IGNORE_AR):
    lib = _abi.load()
    crops = (_abi.Crop * len(rects))()
    for i, (x, y, w, h) in enumerate(rects):
        crops[i].data, crops[i].width, crops[i].height, crops[i].pitch = 0x10000 + y * pitch + 3 * x, w, h, pitch
    p = util.make_pipeline(dsize, util.OPS_C2, aspect=aspect, out_ptr=0x1000)
    out = (C.c_int64 * 12)()
    n_planes = len(rects) if n_planes is None else n_planes
    _abi.check(lib.cvgs_b200_debug_plan(crops, n_planes, len(rects), C.byref(p), sms, image_mode, items_per_warp, out))
    keys = ["ok", "NPB", "HP", "tiles_x", "items", "slot_bytes", "slots", "resident", "grid", "rb_need", "covered", "tiled"]
    return dict(zip(keys, list(out)))


@pytest.mark.parametrize("dsize", [(64, 128), (224, 224), (1, 1), (33, 7), (300, 41), (1000, 3), (129, 257)])
@pytest.mark.parametrize("n", [1, 2, 50, 256, 1000])
@pytest.mark.parametrize("ipw", [1, 4])
def test_item_ranges_tile_the_launch(dsize, n, ipw):
    rng = np.random.default_rng(n + dsize[0])
    rects = [(0, 0, int(rng.integers(1, 1921)), int(rng.integers(1, 1081))) for _ in range(n)]
    for r in rects:  # keep the horizontal scale inside what one 2 KB box can stage
        assert r[2] >= 1
    g = _plan(rects, dsize, items_per_warp=ipw)
    if not g["ok"]:
        pytest.skip("geometry goes to the direct-gather kernel")
    W, H = dsize
    assert g["HP"] == (H + 1) // 2 and g["tiles_x"] == -(-W // (32 * g["NPB"]))
    assert g["items"] == n * g["HP"] * g["tiles_x"]
    assert g["covered"] == g["items"] and g["tiled"] == 1
    assert 1 <= g["slots"] <= 4 and 1 <= g["resident"] <= 5
    assert g["slot_bytes"] >= 128 + 4 * g["rb_need"] and g["slot_bytes"] % 128 == 0
    assert 4 * g["slots"] * g["slot_bytes"] * g["resident"] <= 227 * 1024
    assert g["grid"] <= g["resident"] * 148 and g["grid"] * 4 * ipw <= g["items"] + 4 * ipw


def test_band_width_follows_the_largest_downscale():
    assert _plan([(0, 0, 60, 120)] * 50, (64, 128))["NPB"] == 2          # 64 columns: one band of two groups
    assert _plan([(0, 0, 896, 896)] * 8, (224, 224))["NPB"] == 4         # 4x down-scale of 128 columns fits a 2 KB box
    wide = _plan([(0, 0, 1920, 1080)] * 4, (128, 64))                    # 15x: the band narrows until its span fits
    assert wide["ok"] == 1 and wide["NPB"] == 1
    assert _plan([(0, 0, 1920, 1080)], (16, 16))["ok"] == 0              # 120x: no box is wide enough -> direct kernel
    assert _plan([(0, 0, 100, 100)], (64, 64), pitch=1001)["ok"] == 0    # pitch not a multiple of 16 bytes


def test_small_launch_spreads_or_packs_by_mode():
    rects = [(0, 0, 140, 280)] * 50
    spread, packed = _plan(rects, (64, 128), items_per_warp=1), _plan(rects, (64, 128), items_per_warp=4)
    assert spread["items"] == packed["items"] == 3200
    assert spread["grid"] == min(800, spread["resident"] * 148) and packed["grid"] == 200


def test_overlap_bookkeeping():
    lib = _abi.load()
    q = lambda key, o, s: lib.cvgs_b200_debug_overlap_query(key, o[0], o[1], s[0], s[1])  # noqa: E731
    prev = lib.cvgs_b200_set_overlap(0)
    try:
        assert q(0x10, (0, 100), (1000, 2000)) == 1                      # overlap off: always plain stream order
        lib.cvgs_b200_set_overlap(1)
        key = 0x7770
        assert q(key, (0, 100), (1000, 2000)) == 1                       # first launch seen on a stream: unknown past
        assert q(key, (100, 200), (1000, 2000)) == 0                     # disjoint output, same source: independent
        assert q(key, (150, 250), (1000, 2000)) == 1                     # write-after-write on [150, 200)
        assert q(key, (300, 400), (160, 170)) == 1                       # reads what the previous launch writes
        assert q(key, (165, 180), (5000, 6000)) == 1                     # writes what the launch before it reads
        for i in range(7):                                               # 7 more independent launches fill the window
            assert q(key, (10_000 + 100 * i, 10_100 + 100 * i), (5000, 6000)) == 0
        assert q(key, (20_000, 20_100), (5000, 6000)) == 1               # window of 8 full: forced wait
        assert q(key, (30_000, 30_100), (5000, 6000)) == 0
        assert q(0x8880, (30_000, 30_100), (5000, 6000)) == 1            # another stream has its own history
    finally:
        lib.cvgs_b200_set_overlap(prev)
