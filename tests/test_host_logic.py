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
        assert q(0x20, (0, 100), (1000, 2000)) == 1 and q(0x20, (100, 200), (1000, 2000)) == 1  # mode 1: individual launches keep stream order
        lib.cvgs_b200_set_overlap(2)                                     # mode 2: the caller vouches for what sits between launches
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


# ---------------------------------------------------------------------------------------------------------------
# op-chain normalisation (csrc/preproc_host.hpp build_program) through cvgs_b200_debug_program: no device needed
# ---------------------------------------------------------------------------------------------------------------
def _program(ops, **kw):
    import ctypes as C
    lib = _abi.load()
    out = (C.c_float * 80)()
    p = util.make_pipeline((8, 8), ops, out_ptr=16, **kw)
    rc = lib.cvgs_b200_debug_program(C.byref(p), out)
    if rc != 0:
        raise _abi.CvgsError(lib.cvgs_b200_last_error().decode())
    v = list(out)
    n = int(v[0])
    ops_out = [(int(v[8 + 9 * i]), v[9 + 9 * i:13 + 9 * i], v[13 + 9 * i:17 + 9 * i]) for i in range(n)]
    return dict(n=n, nc_out=int(v[1]), nregs=int(v[2]), special=int(v[3]), dst_chan=[int(x) for x in v[4:8]], ops=ops_out)


MUL, ADD, DIV, FMA, SET, GRAY = 1, 2, 3, 4, 5, 6


def test_program_mul_sub_contracts_and_reorder_becomes_store_offsets():
    g = _program(util.OPS_C2)  # reorder(2,1,0), mul, sub, div
    assert g["n"] == 2 and [k for k, _, _ in g["ops"]] == [FMA, DIV] and not g["special"]
    assert g["dst_chan"][:3] == [2, 1, 0]                      # register (source channel) r is stored as channel 2 - r
    f32 = np.float32
    # constants follow the registers: logical channel 0 (mul 0.3, sub 1.0, div 3.2) lives in register 2
    assert g["ops"][0][1][2] == f32(0.3) and g["ops"][0][2][2] == f32(-1.0) and g["ops"][1][1][2] == f32(3.2)
    assert g["ops"][0][2][0] == f32(-3.2) and g["ops"][1][1][0] == f32(11.8)
    sep = _program(util.OPS_C2, fp_contract=_abi.FP_SEPARATE)
    assert [k for k, _, _ in sep["ops"]] == [MUL, ADD, DIV]


def test_program_alpha_is_hoisted_and_keeps_the_contraction():
    g = _program([("mul", (0.5, 0.25, 2.0)), ("add_alpha", (255.0,)), ("sub", (1.0, 2.0, 3.0, 4.0))])
    assert [k for k, _, _ in g["ops"]] == [SET, FMA] and g["nc_out"] == 4 and g["nregs"] == 4 and g["special"]
    assert g["ops"][0][1][3] == 255.0 and g["ops"][0][2][3] == 1.0           # SET writes register 3 only
    assert g["ops"][0][2][:3] == [0.0, 0.0, 0.0]
    assert g["ops"][1][1][3] == 1.0 and g["ops"][1][2][3] == -4.0            # alpha: 255 * 1 + (-4)
    assert g["dst_chan"] == [0, 1, 2, 3]


def test_program_drop_and_gray():
    g = _program([("reorder", (2, 1, 0, 3)), ("drop_alpha", ()), ("mul", (2.0, 3.0, 4.0))], src_type=_abi.CVGS_8UC4)
    assert g["nc_out"] == 3 and g["nregs"] == 4 and g["special"] and g["dst_chan"] == [2, 1, 0, -1]
    assert g["ops"][0][1] == [4.0, 3.0, 2.0, 1.0]                            # the dropped register gets the identity
    g = _program([("reorder", (2, 1, 0)), ("gray", (0,)), ("sub", (0.5,))])
    kind = g["ops"][0][0]
    assert kind & 0xFF == GRAY and (kind >> 8) & 3 == 2 and (kind >> 12) & 3 == 1 and (kind >> 16) & 3 == 0
    assert not (kind >> 20) & 1 and not (kind >> 21) & 1
    assert g["nc_out"] == 1 and g["dst_chan"] == [0, -1, -1, -1] and g["ops"][1][0] == ADD and g["ops"][1][1][0] == -0.5
    assert (_program([("gray", (1,))], fp_contract=_abi.FP_SEPARATE)["ops"][0][0] >> 20) & 3 == 3


def test_program_rejects_inconsistent_chains():
    for ops, kw in [([("drop_alpha", ())], {}), ([("add_alpha", (1.0,))], dict(src_type=_abi.CVGS_8UC4)),
                    ([("gray", ()), ("gray", ())], {}), ([("add_alpha", (1.0,)), ("add_alpha", (1.0,))], {}),
                    ([("gray", (2,))], {}), ([("reorder", (0, 0, 1))], {})]:
        with pytest.raises(_abi.CvgsError):
            _program(ops, **kw)


def test_fast_division_by_launch_constants_is_exact():
    """The warps' prologues divide item and cost indices (< 2^31) by the plan's constants with multiply-shift (FastDiv,
    csrc/preproc_tma.cuh); the per-warp item ranges tile the launch only if that equals integer division."""
    lib = _abi.load()
    rng = np.random.default_rng(12)
    divisors = [1, 2, 3, 4, 5, 7, 64, 112, 113, 224, 225, 1000, 57344, 65535, 65536, 2 ** 20 + 1, 2 ** 30, 2 ** 31 - 1]
    divisors += [int(v) for v in rng.integers(1, 2 ** 31 - 1, size=60)] + [int(v) for v in rng.integers(1, 5000, size=60)]
    for d in divisors:
        ns = [0, 1, d - 1, d, d + 1, 2 * d - 1, 2 * d, 2 ** 31 - 1, (2 ** 31 - 1) // d * d, max(0, (2 ** 31 - 1) // d * d - 1)]
        ns += [int(v) for v in rng.integers(0, 2 ** 31 - 1, size=200)]
        for n in ns:
            if 0 <= n < 2 ** 31:
                assert lib.cvgs_b200_debug_fast_div(n, d) == n // d, (n, d)


def test_warp_mode_of_the_fast_kernel():
    """Host side of the fast warp kernel: the reciprocal of the perspective denominator skips its range check only when
    every denominator over the (padded) destination is finite, of one sign and comfortably inside the normal range."""
    import ctypes as C
    lib = _abi.load()
    f9 = lambda v: (C.c_float * 9)(*v)  # noqa: E731
    ident = [1, 0, 0, 0, 1, 0, 0, 0, 1]
    assert lib.cvgs_b200_debug_warp_mode(f9(ident), 0, 640, 480) == 0          # affine
    assert lib.cvgs_b200_debug_warp_mode(f9(ident), 1, 640, 480) == 1          # denominator 1 everywhere
    assert lib.cvgs_b200_debug_warp_mode(f9([1, 0, 0, 0, 1, 0, 1e-4, -2e-4, 1]), 1, 640, 480) == 1
    assert lib.cvgs_b200_debug_warp_mode(f9([1, 0, 0, 0, 1, 0, -1e-3, 0, -0.5]), 1, 640, 480) == 1   # negative throughout
    assert lib.cvgs_b200_debug_warp_mode(f9([1, 0, 0, 0, 1, 0, -1e-2, 0, 1]), 1, 640, 480) == 2      # crosses zero at x = 100
    assert lib.cvgs_b200_debug_warp_mode(f9([1, 0, 0, 0, 1, 0, -1e-2, 0, 1]), 1, 64, 480) == 2       # ... also inside the padded width (128)
    assert lib.cvgs_b200_debug_warp_mode(f9([1, 0, 0, 0, 1, 0, -1e-3, 0, 1]), 1, 640, 480) == 1      # zero at x = 1000: outside 640 (padded) columns
    assert lib.cvgs_b200_debug_warp_mode(f9([1, 0, 0, 0, 1, 0, 0, 0, 0]), 1, 640, 480) == 2          # 0 / 0
    assert lib.cvgs_b200_debug_warp_mode(f9([1, 0, 0, 0, 1, 0, 0, 0, 1e-35]), 1, 640, 480) == 2      # too small
    assert lib.cvgs_b200_debug_warp_mode(f9([1, 0, 0, 0, 1, 0, 0, 0, 1e35]), 1, 640, 480) == 2       # too large
    assert lib.cvgs_b200_debug_warp_mode(f9([1, 0, 0, 0, 1, 0, float("nan"), 0, 1]), 1, 640, 480) == 2
    assert lib.cvgs_b200_debug_warp_mode(f9([1, 0, 0, 0, 1, 0, 1.0, 1.0, 1e-9]), 1, 640, 480) == 2   # near-cancellation at (0, 0): margin
    assert lib.cvgs_b200_debug_warp_mode(None, 1, 640, 480) == -1


def test_chain_kinds_of_the_tma_kernel():
    """Which instantiation a chain runs on (scaled_program): no op and lone linear ops are the specialised shape with a
    division by 1, rounded interpolation and long chains the interpreter, gray / alpha conversions their own kinds; the
    alpha plane's value is the constant run through the ops behind the conversion in IEEE single precision."""
    import ctypes as C
    import numpy as np
    lib = _abi.load()

    def kind(ops, **kw):
        p = util.make_pipeline((64, 64), ops, out_ptr=0x1000, **kw)
        a = C.c_float(0)
        return lib.cvgs_b200_debug_chain_kind(C.byref(p), C.byref(a)), a.value

    assert kind([])[0] == 1
    assert kind([("mul", (0.5, 0.25, 2.0))])[0] == 1
    assert kind([("mul", (0.5, 0.25, 2.0)), ("sub", (1.0, 2.0, 3.0))])[0] == 1          # contracted into one FMA
    assert kind([("mul", (1 / 255.0,) * 3), ("sub", (0.485, 0.456, 0.406)), ("div", (0.229, 0.224, 0.225))])[0] == 1
    assert kind([("div", (3.0, 5.0, 7.0))])[0] == 1
    assert kind([("mul", (0.5, 0.25, 2.0)), ("sub", (1.0, 2.0, 3.0))], fp_contract=_abi.FP_SEPARATE)[0] == 0   # two roundings: two ops
    assert kind([("div", (3.0, 5.0, 7.0)), ("mul", (0.5, 0.25, 2.0))])[0] == 0
    assert kind([("mul", (0.5, 0.25, 2.0))], interp_mode=_abi.INTERP_ROUND_U8)[0] == 0
    assert kind([("mul", (1e30, 1.0, 1.0)), ("div", (3.0, 5.0, 7.0))])[0] == 0          # outside the proven range
    assert kind([("gray", (1,)), ("mul", (1 / 255.0,))])[0] == 2
    k, a = kind([("add_alpha", (255.0,))])
    assert (k, a) == (3, 255.0)
    k, a = kind([("reorder", (2, 1, 0)), ("add_alpha", (255.0,)), ("mul", (1 / 255.0,) * 4), ("sub", (0.485, 0.456, 0.406, 0.5)),
                 ("div", (0.229, 0.224, 0.225, 0.25))])
    # mul + sub contract into one FMA (one rounding: the product of two floats is exact in double), then an IEEE division
    want = np.float32(255.0 * float(np.float32(1 / 255.0)) - 0.5)
    want = np.float32(want / np.float32(0.25))
    assert k == 3 and np.float32(a) == want, (k, a, want)
    k, a = kind([("add_alpha", (7.0,)), ("mul", (2.0, 2.0, 2.0, 3.0)), ("add", (1.0, 1.0, 1.0, 0.5)), ("mul", (1.0, 1.0, 1.0, 2.0))])
    assert k == 4 and a == (7.0 * 3.0 + 0.5) * 2.0
