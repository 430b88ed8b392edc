"""CPU tests: pin oracle/oracle.c against (a) the known answers the reference's own tests hold for this
path, (b) an independent numpy restatement, (c) OpenCV-CPU for the stages where
OpenCV-CPU and OpenCV-CUDA agree (convertTo / subtract / divide / split), (d) the golden vectors that
were produced by running the reference's real fused kernel on a B200 (tests/golden/)."""
import ctypes as C
import glob
import os

import numpy as np
import pytest

from cvgpuspeedup_b200 import _abi
from tests import np_restatement as npr
from tests import util

HERE = os.path.dirname(os.path.abspath(__file__))


def test_constant_image_known_answer():
    """tests/batchresize/test_batchresize_x_split3D.cu:58-62: image = (5,5,5), crops Rect(i,i,60,120)
    -> 64x128, RGB2BGR, *0.3, -(1,4,3.2), /(3.2,0.6,11.8).  Interpolating a constant is exact up to the
    float weights, so the reference checks |a-b| <= 1e-4 (tests/testsCommon.cuh:36-41)."""
    img = np.full((2160, 3 * 3840), 5, dtype=np.uint8)
    rects = [(i, i, 60, 120) for i in range(10)]
    out = util.run_oracle(img, rects, (64, 128), util.OPS_C2)
    sub, div = (1.0, 4.0, 3.2), (3.2, 0.6, 11.8)
    for c in range(3):
        expect = (np.float32(5) * np.float32(0.3) - np.float32(sub[c])) / np.float32(div[c])
        assert np.all(np.abs(out[:, c] - expect) <= 1e-4)


def test_split_known_answer():
    """tests/unit_tests/test_split.cu:47-90: a (1,2,3) 16x16 image splits into planes 1, 2, 3 (exact)."""
    img = np.tile(np.array([1, 2, 3], dtype=np.uint8), (16, 16))
    out = util.run_oracle(img, [(0, 0, 16, 16)] * 10, (16, 16), [])
    for c in range(3):
        assert np.all(out[:, c] == c + 1)


def test_convert_to_known_answer():
    """tests/single_operation/test_convertTo.cu:60-73: (20,30,40) * 0.5 + 0.5."""
    img = np.tile(np.array([20, 30, 40], dtype=np.uint8), (8, 8))
    out = util.run_oracle(img, [(0, 0, 8, 8)], (8, 8), [("mul", (0.5,) * 3), ("add", (0.5,) * 3)],
                          fp_contract=_abi.FP_SEPARATE)
    assert np.all(out[0, 0] == 10.5) and np.all(out[0, 1] == 15.5) and np.all(out[0, 2] == 20.5)


def test_active_threads_extent():
    """fkl/tests/algorithm/test_crop.cu:39-45: batch of 2 rects resized to 100x100 -> (100,100,2) elements,
    all written."""
    rng = np.random.default_rng(0)
    img = util.make_image(rng, 300, 200)
    out = util.run_oracle(img, [(0, 0, 34, 25), (40, 40, 70, 15)], (100, 100), [])
    assert out.shape == (2, 3, 100, 100) and not np.isnan(out).any()


@pytest.mark.parametrize("order,expect", [(_abi.CT_NEWEST_FIRST, lambda z: 100 - z),
                                          (_abi.CT_OLDEST_FIRST, lambda z: 100 - (15 - z - 1))])
@pytest.mark.parametrize("mode", [_abi.CT_STANDARD, _abi.CT_TRANSPOSED])
def test_circular_tensor_known_answer(order, expect, mode):
    """tests/batchread/test_circularbatchread_x_write3D.cu:211,266,391,448: after 100 updates with frame
    value i+1, plane z holds 100-z (NewestFirst) / 100-(15-z-1) (OldestFirst)."""
    lib = util.oracle_lib()
    B, W, H = 15, 32, 24
    t = lib.oracle_ct_create(W, H, 3, B, order, mode)
    p = util.make_pipeline((W, H), [])
    for i in range(100):
        img = np.full((H, 3 * W), i + 1, dtype=np.uint8)
        crop = util.host_crops(img, [(0, 0, W, H)])
        assert lib.oracle_ct_update(t, crop, C.byref(p), 1) == 0
    data = np.ctypeslib.as_array(lib.oracle_ct_data(t), shape=(B * 3 * H * W,)).copy()
    lib.oracle_ct_destroy(t)
    data = data.reshape(B, 3, H, W) if mode == _abi.CT_STANDARD else data.reshape(3, B, H, W).swapaxes(0, 1)
    for z in range(B):
        assert np.all(data[z] == expect(z)), z


def test_circular_batch_rotation_index():
    """test_circularbatchread_x_write3D.cu:77-81: Ascendent read with first=4 of 15 maps z -> (z+4) mod 15."""
    lib = util.oracle_lib()
    B, W, H = 15, 8, 8
    t = lib.oracle_ct_create(W, H, 3, B, _abi.CT_OLDEST_FIRST, _abi.CT_STANDARD)
    p = util.make_pipeline((W, H), [])
    for i in range(B + 5):  # ring full; the last update runs with m_nextUpdateIdx == 4
        img = np.full((H, 3 * W), i, dtype=np.uint8)
        lib.oracle_ct_update(t, util.host_crops(img, [(0, 0, W, H)]), C.byref(p), 1)
    tmp = np.ctypeslib.as_array(lib.oracle_ct_temp(t), shape=(B, 3, H, W)).copy()
    pub = np.ctypeslib.as_array(lib.oracle_ct_data(t), shape=(B, 3, H, W)).copy()
    lib.oracle_ct_destroy(t)
    for z in range(B):
        assert np.all(pub[z] == tmp[(z + 4) % B])


@pytest.mark.parametrize("sw,sh,dw,dh", [(60, 120, 64, 128), (30, 120, 64, 128), (640, 480, 64, 128),
                                         (1000, 37, 224, 224), (37, 1000, 224, 224), (7, 5, 64, 128)])
@pytest.mark.parametrize("aspect", [0, 1, 2, 3])
def test_geometry_matches_restatement(sw, sh, dw, dh, aspect):
    lib = util.oracle_lib()
    g = lib.Geom()
    lib.oracle_resize_geometry(sw, sh, dw, dh, aspect, C.byref(g))
    fx, fy, x1, y1, x2, y2 = npr.geometry(sw, sh, dw, dh, aspect)
    assert (np.float32(g.fx), np.float32(g.fy), g.x1, g.y1, g.x2, g.y2) == (fx, fy, x1, y1, x2, y2)
    assert 0 <= g.x1 <= g.x2 < dw and 0 <= g.y1 <= g.y2 < dh


def test_preserve_ar_benchmark_shape():
    """benchmarks/benchmark_CPUandGPU_cvGS_vs_fk.cu:55-72: 30x120 crops into 64x128 keep a 32-wide
    centred band; the rest is chain(background)."""
    img = np.full((400, 3 * 400), 5, dtype=np.uint8)
    out = util.run_oracle(img, [(3, 3, 30, 120)], (64, 128), [], aspect=_abi.PRESERVE_AR,
                          background=(128.0, 128.0, 128.0))
    assert np.all(out[0, :, :, :16] == 128) and np.all(out[0, :, :, 48:] == 128)
    assert np.all(np.abs(out[0, :, :, 16:48] - 5) <= 1e-4)


@pytest.mark.parametrize("fused", [True, False])
@pytest.mark.parametrize("aspect", [1, 0, 2, 3])
def test_oracle_equals_numpy_restatement(fused, aspect):
    rng = np.random.default_rng(11)
    img = util.make_image(rng, 320, 200, pitch=1024)
    rects = [(0, 0, 320, 200), (5, 7, 24, 48), (100, 3, 199, 33), (17, 150, 7, 5), (300, 0, 20, 200), (1, 1, 64, 128)]
    ops = util.OPS_C2
    bg = (7.0, 128.0, 250.5)
    got = util.run_oracle(img, rects, (64, 128), ops, aspect=aspect, background=bg, n_planes=8,
                          fp_contract=_abi.FP_REFERENCE_FUSED if fused else _abi.FP_SEPARATE)
    want = npr.preproc(img, 320, rects, (64, 128), ops, aspect=aspect, bg=bg, n_planes=8, fused=fused)
    util.assert_bit_equal(got, want, "oracle.c vs numpy restatement")


def test_oracle_round_u8_and_layouts():
    rng = np.random.default_rng(12)
    img = util.make_image(rng, 128, 96)
    rects = [(0, 0, 128, 96), (10, 10, 50, 60)]
    ops = util.OPS_C1
    want = npr.preproc(img, 128, rects, (40, 56), ops, fused=False, round_u8=True)
    nchw = util.run_oracle(img, rects, (40, 56), ops, fp_contract=_abi.FP_SEPARATE, interp_mode=_abi.INTERP_ROUND_U8)
    util.assert_bit_equal(nchw, want, "round_u8")
    cnhw = util.run_oracle(img, rects, (40, 56), ops, fp_contract=_abi.FP_SEPARATE, interp_mode=_abi.INTERP_ROUND_U8,
                           layout=_abi.OUT_CNHW)
    util.assert_bit_equal(cnhw, want.swapaxes(0, 1), "CNHW")
    nhwc = util.run_oracle(img, rects, (40, 56), ops, fp_contract=_abi.FP_SEPARATE, interp_mode=_abi.INTERP_ROUND_U8,
                           layout=_abi.OUT_NHWC)
    util.assert_bit_equal(nhwc, np.moveaxis(want, 1, -1), "NHWC")


def test_oracle_split_write_equals_tensor_split():
    """fk::SplitWrite (one image per crop and channel, own pitch) holds the same values as TensorSplit."""
    import ctypes as C
    w = util.workload_c2(seed=8, n=4, frame=(640, 600), pitch=1920)
    W, H = w.dsize
    ref = util.run_oracle(w.image, w.rects, w.dsize, w.ops)
    bufs = [[np.full((H, W + 5 * c), np.nan, dtype=np.float32) for c in range(3)] for _ in w.rects]
    arr = (_abi.Plane * 12)()
    for z in range(4):
        for c in range(3):
            arr[3 * z + c].data, arr[3 * z + c].pitch_bytes = bufs[z][c].ctypes.data, bufs[z][c].strides[0]
    p = util.make_pipeline(w.dsize, w.ops, layout=_abi.OUT_PLANES, out_ptr=C.addressof(arr))
    assert util.oracle_lib().oracle_preproc(util.host_crops(w.image, w.rects), 4, 4, C.byref(p), 0) == 0
    for z in range(4):
        for c in range(3):
            util.assert_bit_equal(bufs[z][c][:, :W], ref[z, c], f"plane {z} channel {c}")
            assert np.isnan(bufs[z][c][:, W:]).all()


@pytest.mark.parametrize("src_type,dtype,nc", [(_abi.CVGS_16UC3, np.uint16, 3), (_abi.CVGS_16SC3, np.int16, 3),
                                               (_abi.CVGS_8UC4, np.uint8, 4), (_abi.CVGS_16UC4, np.uint16, 4),
                                               (_abi.CVGS_16SC4, np.int16, 4)])
def test_oracle_16bit_sources_equal_numpy(src_type, dtype, nc):
    """ushort3 / short3 sources: same index math and rounding order on exactly converted taps (independent numpy
    restatement of interpolation.cuh:57-92 with float32 arithmetic and explicit fma emulation in float64)."""
    rng = np.random.default_rng(12)
    w, h, W, H = 37, 29, 16, 24
    pb = nc * np.dtype(dtype).itemsize
    img = rng.integers(0, 256, size=(h, pb * w + 10), dtype=np.uint8)
    px = img[:, :pb * w].copy().view(dtype).reshape(h, w, nc).astype(np.float32)
    out = util.run_oracle(img, [(0, 0, w, h)], (W, H), [], src_type=src_type)[0]
    fx, fy = np.float32(1.0 / (W / w)), np.float32(1.0 / (H / h))

    def fma(a, b, c):  # float32 fma through float64: exact product (24x24 bits), one rounding of the sum
        return np.float32(np.float64(a) * np.float64(b) + np.float64(c))
    for y in range(H):
        for x in range(W):
            sx, sy = np.float32(x) * fx, np.float32(y) * fy
            x1, y1 = int(np.floor(sx)), int(np.floor(sy))
            x2r, y2r = min(x1 + 1, w - 1), min(y1 + 1, h - 1)
            wx1, wx0 = sx - np.float32(x1), np.float32(x1 + 1) - sx
            wy1, wy0 = sy - np.float32(y1), np.float32(y1 + 1) - sy
            w00, w10, w01, w11 = wx0 * wy0, wx1 * wy0, wx0 * wy1, wx1 * wy1
            for c in range(nc):
                t = px[y1, x2r, c] * w10
                t = fma(px[y1, x1, c], w00, t)
                t = fma(px[y2r, x1, c], w01, t)
                t = fma(px[y2r, x2r, c], w11, t)
                assert out[c, y, x] == t, (x, y, c)


def test_post_resize_stages_match_opencv_cpu():
    """BASELINE config 1, read per SURVEY F3: OpenCV-CPU is an exact oracle for convertTo / subtract /
    divide / split, not for the resize stage.  Identity-size 'resize' isolates those stages."""
    cv2 = pytest.importorskip("cv2")
    w = util.workload_c1()
    src = w.image.reshape(480, 640, 3)
    got = util.run_oracle(w.image, w.rects, (640, 480), util.OPS_C1, fp_contract=_abi.FP_SEPARATE)
    f = src.astype(np.float32)
    f = cv2.multiply(f, (0.5, 0.5, 0.5, 0))
    f = cv2.subtract(f, (1.0, 4.0, 6.0, 0))
    f = cv2.divide(f, (2.0, 8.0, 1.0, 1.0))
    planes = cv2.split(f)
    for c in range(3):
        assert util.ulp_diff(got[0, c], planes[c]).max() <= 1, c  # 1 ULP, BASELINE.json north_star


def test_golden_vectors_from_reference_kernel():
    """Golden vectors = outputs of the reference's own fused kernel (oracle/_ref/libfkref_16.so) captured
    on a B200 by tests/golden/make_golden.py.  The oracle must reproduce them bit for bit."""
    files = sorted(glob.glob(os.path.join(HERE, "golden", "*.npz")))
    if not files:
        pytest.skip("golden vectors not generated yet (tests/golden/make_golden.py on a GPU box)")
    for f in files:
        g = np.load(f)
        if "warp_type" in g or "code" in g:  # warp / colour-conversion vectors: tests/test_warp_cpu.py, test_cvtcolor_cpu.py
            continue
        if "standard" in g:  # NV12 frame: ReadYUV + ConvertYUVToRGB + resize + mul/sub/div of the reference
            import ctypes as C
            img = np.ascontiguousarray(g["image"])
            crop = (_abi.Crop * 1)()
            crop[0].data, crop[0].width, crop[0].height, crop[0].pitch = img.ctypes.data, int(g["width"]), int(g["height"]), img.shape[1]
            dsize = tuple(int(v) for v in g["dsize"])
            got = np.full((1, 3, dsize[1], dsize[0]), np.nan, dtype=np.float32)
            p = util.make_pipeline(dsize, [("mul", tuple(g["mul"])), ("sub", tuple(g["sub"])), ("div", tuple(g["div"]))],
                                   out_ptr=got.ctypes.data, src_type=int(g["src_type"]) if "src_type" in g else _abi.CVGS_NV12,
                                   yuv_standard=int(g["standard"]))
            assert util.oracle_lib().oracle_preproc(crop, 1, 1, C.byref(p), 0) == 0
            util.assert_bit_equal(got[0], g["out"], os.path.basename(f))
            continue
        src_type = int(g["src_type"]) if "src_type" in g else _abi.CVGS_8UC3
        nc = util.channels_of(src_type)
        rects = [tuple(int(v) for v in r) for r in g["rects"]]
        ops = []
        if int(g["swap"]):
            ops.append(("reorder", (2, 1, 0) if nc == 3 else (2, 1, 0, 3)))
        ops += [("mul", tuple(g["mul"])), ("sub", tuple(g["sub"])), ("div", tuple(g["div"]))]
        got = util.run_oracle(g["image"], rects, tuple(int(v) for v in g["dsize"]), ops, aspect=int(g["aspect"]),
                              background=tuple(float(v) for v in g["bg"]), n_planes=int(g["n_planes"]),
                              used=int(g["used"]), src_type=src_type)
        util.assert_bit_equal(got, g["out"], os.path.basename(f))


def test_round_to_source_depth_of_small_negative_values_is_plus_zero():
    """CVGS_INTERP_ROUND_U8 on CV_16SC3: the interpolated value passes through a short, so (-0.5, -0] comes back as +0
    (found by the fuzz soak: the oracle used to return the -0 of nearbyintf)."""
    img = np.zeros((2, 16), dtype=np.int16)
    img[0, :3] = -1          # pixel (0, 0) = -1, pixel (1, 0) = 0: a 2x up-scale interpolates -0.5 and -0.25... in between
    crops_img = img.view(np.uint8)
    out = util.run_oracle(crops_img, [(0, 0, 2, 2)], (8, 8), [], src_type=_abi.CVGS_16SC3, interp_mode=_abi.INTERP_ROUND_U8)
    assert not np.signbit(out[out == 0]).any()
    assert (out[0, :, 0, 0] == -1).all() and (out == 0).any()
