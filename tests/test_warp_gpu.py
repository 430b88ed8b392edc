"""Batched affine / perspective warp in front of the chain (cvGS::warp, reference include/cvGPUSpeedup.cuh:266-442 ->
fk::Warping, fkl/.../image_processing/warping.cuh:43-91).  Bit-exact against the reference's own kernel (one image per
call: its batch size is a template parameter) and the oracle, on the matrices of the reference's test
(tests/warping/test_warping_opencv.cu:49-51,97-100,137-150) and random ones."""
import ctypes as C
import math

import numpy as np
import pytest
import torch

import cvgpuspeedup_b200 as cvgs
from cvgpuspeedup_b200 import _abi
from tests import gpu_util, util

pytestmark = pytest.mark.gpu


def perspective_from_points(src, dst):
    """cv::getPerspectiveTransform: the 3x3 H (h22 = 1) with H * src_i ~ dst_i, solved in double."""
    a, b = [], []
    for (x, y), (u, v) in zip(src, dst):
        a.append([x, y, 1, 0, 0, 0, -x * u, -y * u]); b.append(u)
        a.append([0, 0, 0, x, y, 1, -x * v, -y * v]); b.append(v)
    h = np.linalg.solve(np.array(a, dtype=np.float64), np.array(b, dtype=np.float64))
    return np.append(h, 1.0).reshape(3, 3)


REF_SRC = [(56, 65), (368, 52), (28, 387), (389, 390)]
REF_DST = [(0, 0), (300, 0), (0, 300), (300, 300)]


def _matrices(rng, n, warp_type, w, h):
    out = []
    for i in range(n):
        if warp_type == cvgs.WARP_AFFINE:
            ang, sc = rng.uniform(-math.pi, math.pi), rng.uniform(0.4, 2.5)
            m = np.array([[sc * math.cos(ang), -sc * math.sin(ang), rng.uniform(-w / 2, w / 2)],
                          [sc * math.sin(ang), sc * math.cos(ang), rng.uniform(-h / 2, h / 2)]])
        else:
            jit = lambda: rng.uniform(-0.2, 0.2)  # noqa: E731
            src = [(w * (0.1 + jit()), h * (0.1 + jit())), (w * (0.9 + jit()), h * (0.1 + jit())),
                   (w * (0.1 + jit()), h * (0.9 + jit())), (w * (0.9 + jit()), h * (0.9 + jit()))]
            dst = [(0, 0), (w * 0.8, 0), (0, h * 0.8), (w * 0.8, h * 0.8)]
            m = perspective_from_points(src, dst)
        out.append(m)
    return out


def _launch(d_imgs, sizes, pitch, inverses, warp_type, dsize, ops, n_planes=None, used=None, background=(0, 0, 0),
            layout=_abi.OUT_NCHW, u8=None, src_type=_abi.CVGS_8UC3):
    lib = _abi.load()
    n = len(d_imgs)
    n_planes = n if n_planes is None else n_planes
    used = n if used is None else used
    crops = (_abi.Crop * n)()
    warps = (_abi.Warp * n)()
    for i, (t, (w, h)) in enumerate(zip(d_imgs, sizes)):
        crops[i].data, crops[i].width, crops[i].height, crops[i].pitch = t.data_ptr(), w, h, pitch
        warps[i].type = warp_type
        for k in range(9):
            warps[i].m[k] = float(inverses[i][k])
    if u8 is None:
        out = torch.full(util.out_shape(n_planes, dsize, layout, 0, util.channels_of(src_type)), float("nan"), device="cuda")
        p = util.make_pipeline(dsize, ops, out_ptr=out.data_ptr(), layout=layout, background=background, src_type=src_type)
    else:
        out = torch.full((n_planes, dsize[1], dsize[0], 3), 99, dtype=torch.uint8, device="cuda")
        p = util.make_pipeline(dsize, ops, out_ptr=out.data_ptr(), layout=_abi.OUT_NHWC, background=background,
                               dst_type=_abi.CVGS_8UC3, u8_cast=u8)
    _abi.check(lib.cvgs_b200_warp_launch(crops, warps, n_planes, used, C.byref(p), None))
    torch.cuda.synchronize()
    return out.cpu().numpy()


def _oracle(imgs, sizes, pitch, inverses, warp_type, dsize, ops, n_planes=None, used=None, background=(0, 0, 0),
            layout=_abi.OUT_NCHW, u8=None, src_type=_abi.CVGS_8UC3):
    n = len(imgs)
    n_planes = n if n_planes is None else n_planes
    used = n if used is None else used
    crops = (_abi.Crop * n)()
    warps = (_abi.Warp * n)()
    for i, (im, (w, h)) in enumerate(zip(imgs, sizes)):
        crops[i].data, crops[i].width, crops[i].height, crops[i].pitch = im.ctypes.data, w, h, pitch
        warps[i].type = warp_type
        for k in range(9):
            warps[i].m[k] = float(inverses[i][k])
    if u8 is None:
        out = np.full(util.out_shape(n_planes, dsize, layout, 0, util.channels_of(src_type)), np.nan, dtype=np.float32)
        p = util.make_pipeline(dsize, ops, out_ptr=out.ctypes.data, layout=layout, background=background, src_type=src_type)
    else:
        out = np.full((n_planes, dsize[1], dsize[0], 3), 99, dtype=np.uint8)
        p = util.make_pipeline(dsize, ops, out_ptr=out.ctypes.data, layout=_abi.OUT_NHWC, background=background,
                               dst_type=_abi.CVGS_8UC3, u8_cast=u8)
    assert util.oracle_lib().oracle_warp(crops, warps, n_planes, used, C.byref(p), 0) == 0
    return out


def _reference_cases():
    w, h = 420, 410
    persp = perspective_from_points(REF_SRC, REF_DST)
    persp2 = perspective_from_points(REF_SRC, [(0, 0), (200, 0), (0, 200), (200, 200)])
    return w, h, [
        (cvgs.WARP_AFFINE, np.array([[1, 0, 50], [0, 1, 100]], dtype=np.float64)),       # testAffine :97-100
        (cvgs.WARP_AFFINE, np.array([[0.8, 0.3, -20.5], [-0.25, 1.1, 33.25]])),
        (cvgs.WARP_PERSPECTIVE, persp),                                                  # testPerspective :49-55
        (cvgs.WARP_PERSPECTIVE, persp2),                                                 # testPerspectiveBatch
    ]


@pytest.mark.parametrize("case", range(4))
def test_warp_matches_reference_kernel_and_oracle(case):
    if gpu_util.fkref_lib(16) is None:
        pytest.skip("oracle/_ref not built")
    w, h, cases = _reference_cases()
    warp_type, m = cases[case]
    rng = np.random.default_rng(100 + case)
    pitch = 1280
    img = util.make_image(rng, w, h, pitch, smooth=(case % 2 == 0))
    d = gpu_util.device_image(img)
    inv = cvgs.api.invert_warp_matrix(m, warp_type)
    mul = (0.5, 1.25, 1 / 255.0)
    for dsize in [(w, h), (300, 300), (123, 77)]:
        ref = gpu_util.run_fkref_warp(img, w, h, warp_type, inv, dsize, mul=mul, d_image=d)
        ours = _launch([d], [(w, h)], pitch, [inv], warp_type, dsize, [("mul", mul)])[0]
        orc = _oracle([img], [(w, h)], pitch, [inv], warp_type, dsize, [("mul", mul)])[0]
        util.assert_bit_equal(ours, ref, f"case {case} dsize {dsize}: ours vs reference kernel")
        util.assert_bit_equal(orc, ref, f"case {case} dsize {dsize}: oracle vs reference kernel")
        if dsize == (w, h):
            assert np.count_nonzero(ref) > ref.size // 10  # the warp lands inside the image
        # fk::Cast<float3, uchar3> + packed write (the chain of the reference's test)
        ref8 = gpu_util.run_fkref_warp(img, w, h, warp_type, inv, dsize, mul=None, d_image=d)
        ours8 = _launch([d], [(w, h)], pitch, [inv], warp_type, dsize, [], u8=1)[0]
        orc8 = _oracle([img], [(w, h)], pitch, [inv], warp_type, dsize, [], u8=1)[0]
        assert np.array_equal(ours8, ref8) and np.array_equal(orc8, ref8)


@pytest.mark.parametrize("warp_type", [cvgs.WARP_AFFINE, cvgs.WARP_PERSPECTIVE])
def test_random_warps_match_reference_kernel(warp_type):
    if gpu_util.fkref_lib(16) is None:
        pytest.skip("oracle/_ref not built")
    rng = np.random.default_rng(7 + warp_type)
    w, h, pitch = 333, 251, 1024
    img = util.make_image(rng, w, h, pitch)
    d = gpu_util.device_image(img)
    for m in _matrices(rng, 6, warp_type, w, h):
        inv = cvgs.api.invert_warp_matrix(m, warp_type)
        dsize = (int(rng.integers(1, 400)), int(rng.integers(1, 300)))
        ref = gpu_util.run_fkref_warp(img, w, h, warp_type, inv, dsize, mul=(1, 1, 1), d_image=d)
        ours = _launch([d], [(w, h)], pitch, [inv], warp_type, dsize, [("mul", (1, 1, 1))])[0]
        util.assert_bit_equal(ours, ref, f"type {warp_type} dsize {dsize}")


@pytest.mark.parametrize("layout", [_abi.OUT_NCHW, _abi.OUT_CNHW, _abi.OUT_NHWC])
@pytest.mark.parametrize("warp_type", [cvgs.WARP_AFFINE, cvgs.WARP_PERSPECTIVE])
def test_batched_warp_with_chain_matches_oracle(layout, warp_type):
    """A batch larger than one launch's descriptor table, unused planes, the full chain, every tensor layout."""
    rng = np.random.default_rng(31 + layout + 10 * warp_type)
    pitch = 768
    sizes = [(int(rng.integers(8, 250)), int(rng.integers(8, 200))) for _ in range(70)]
    imgs = [util.make_image(rng, w, h, pitch) for (w, h) in sizes]
    d = [gpu_util.device_image(im) for im in imgs]
    inverses = [cvgs.api.invert_warp_matrix(m, warp_type)
                for (w, h) in sizes for m in _matrices(rng, 1, warp_type, w, h)]
    dsize = (61, 45)
    kw = dict(n_planes=75, used=70, background=(3.0, 5.0, 7.0), layout=layout)
    ours = _launch(d, sizes, pitch, inverses, warp_type, dsize, util.OPS_C2, **kw)
    orc = _oracle(imgs, sizes, pitch, inverses, warp_type, dsize, util.OPS_C2, **kw)
    util.assert_bit_equal(ours, orc, f"layout {layout} type {warp_type}")


def test_warp_u8_saturating_output_matches_oracle():
    rng = np.random.default_rng(5)
    w, h, pitch = 200, 150, 640
    img = util.make_image(rng, w, h, pitch)
    d = gpu_util.device_image(img)
    inv = cvgs.api.invert_warp_matrix(np.array([[1.3, 0.2, -30], [-0.1, 0.9, 12]]), cvgs.WARP_AFFINE)
    ops = [("mul", (1.7, 0.5, 1.0)), ("sub", (20.0, -3.0, 0.5))]  # drives values outside [0, 255]
    for u8 in (0,):
        ours = _launch([d], [(w, h)], pitch, [inv], cvgs.WARP_AFFINE, (180, 140), ops, u8=u8)
        orc = _oracle([img], [(w, h)], pitch, [inv], cvgs.WARP_AFFINE, (180, 140), ops, u8=u8)
        assert np.array_equal(ours, orc)


def test_identity_warp_returns_the_image():
    rng = np.random.default_rng(9)
    w, h, pitch = 129, 65, 512
    img = util.make_image(rng, w, h, pitch)
    d = gpu_util.device_image(img)
    inv = cvgs.api.invert_warp_matrix(np.array([[1, 0, 0], [0, 1, 0]]), cvgs.WARP_AFFINE)
    ours = _launch([d], [(w, h)], pitch, [inv], cvgs.WARP_AFFINE, (w, h), [], u8=1)[0]
    assert np.array_equal(ours, img[:, :3 * w].reshape(h, w, 3))


def test_python_api_warp_matches_abi_call():
    rng = np.random.default_rng(11)
    w, h, pitch = 160, 120, 512
    img = util.make_image(rng, w, h, pitch)
    t = gpu_util.device_image(img)
    mat = cvgs.GpuMat(t.data_ptr(), w, h, pitch, owner=t)
    m = perspective_from_points([(10, 12), (150, 8), (5, 110), (155, 115)], [(0, 0), (100, 0), (0, 100), (100, 100)])
    out = torch.full((2, 3, 100, 100), float("nan"), device="cuda")
    cvgs.executeOperations(None, cvgs.warp([mat, mat], [m, m], (100, 100), cvgs.WARP_PERSPECTIVE),
                           cvgs.multiply((0.5, 0.5, 0.5)), cvgs.split(out))
    torch.cuda.synchronize()
    inv = cvgs.api.invert_warp_matrix(m, cvgs.WARP_PERSPECTIVE)
    want = _launch([t, t], [(w, h)] * 2, pitch, [inv, inv], cvgs.WARP_PERSPECTIVE, (100, 100), [("mul", (0.5,) * 3)])
    util.assert_bit_equal(out.cpu().numpy(), want, "python api")
    out8 = torch.zeros((100, 100, 3), dtype=torch.uint8, device="cuda")
    cvgs.executeOperations(None, cvgs.warp(mat, m, (100, 100), cvgs.WARP_PERSPECTIVE), cvgs.write_u8(out8, cast=True))
    torch.cuda.synchronize()
    want8 = _launch([t], [(w, h)], pitch, [inv], cvgs.WARP_PERSPECTIVE, (100, 100), [], u8=1)[0]
    assert np.array_equal(out8.cpu().numpy(), want8)


def test_warp_rejects_bad_input():
    lib = _abi.load()
    t = torch.zeros(64 * 64 * 3, dtype=torch.uint8, device="cuda")
    crops = (_abi.Crop * 1)()
    crops[0].data, crops[0].width, crops[0].height, crops[0].pitch = t.data_ptr(), 64, 64, 192
    warps = (_abi.Warp * 1)()
    warps[0].type = 7
    out = torch.zeros(3 * 8 * 8, device="cuda")
    p = util.make_pipeline((8, 8), [], out_ptr=out.data_ptr())
    assert lib.cvgs_b200_warp_launch(crops, warps, 1, 1, C.byref(p), None) != 0
    assert b"bad type" in lib.cvgs_b200_last_error()
    warps[0].type = 0
    pnv = util.make_pipeline((8, 8), [], out_ptr=out.data_ptr(), src_type=_abi.CVGS_NV12)
    assert lib.cvgs_b200_warp_launch(crops, warps, 1, 1, C.byref(pnv), None) != 0
    assert lib.cvgs_b200_warp_launch(None, warps, 1, 1, C.byref(p), None) != 0


@pytest.mark.parametrize("seed", range(16 + int(__import__("os").environ.get("CVGS_FUZZ_EXTRA", "0"))))
def test_wild_matrices_against_oracle(seed):
    """Raw inverse matrices straight into the C-ABI: huge and tiny coefficients, denominators that cross zero inside
    the destination (infinite / NaN coordinates), mirrored and degenerate transforms -- whatever the coordinate is, both
    sides must agree on inside / outside and on every bit of the interpolation."""
    rng = np.random.default_rng(8000 + seed)
    w, h, pitch = int(rng.integers(2, 200)), int(rng.integers(2, 150)), 640
    img = util.make_image(rng, w, h, pitch)
    d = gpu_util.device_image(img)
    warp_type = seed % 2
    n = 5
    inverses = []
    for i in range(n):
        scale = float(10.0 ** rng.uniform(-3, 2))
        m = rng.normal(0, scale, size=9).astype(np.float32)
        m[2], m[5] = rng.uniform(-w, 2 * w), rng.uniform(-h, 2 * h)
        if warp_type == 1:
            m[6], m[7] = rng.normal(0, 0.02, size=2)
            m[8] = rng.choice([1.0, -0.5, 0.0, 1e-30, 3.0])
        if i == 3:
            m[:] = 0.0                       # everything maps to (0, 0) (affine) or to NaN (perspective: 0 / 0)
        if i == 4:
            m[:6] = [1, 0, 0.5, 0, 1, -0.25]  # near identity with fractional shifts
        inverses.append(m)
    dsize = (int(rng.integers(1, 130)), int(rng.integers(1, 90)))
    ops = util.OPS_C2 if seed % 3 else []
    ours = _launch([d] * n, [(w, h)] * n, pitch, inverses, warp_type, dsize, ops)
    orc = _oracle([img] * n, [(w, h)] * n, pitch, inverses, warp_type, dsize, ops)
    util.assert_bit_equal(ours, orc, f"seed {seed} type {warp_type} image {w}x{h} dsize {dsize}")
    # the same launch through the general kernel (variant 1 keeps the chain interpreter and four pixels per lane)
    lib = _abi.load()
    prev = lib.cvgs_b200_set_kernel_variant(1)
    try:
        general = _launch([d] * n, [(w, h)] * n, pitch, inverses, warp_type, dsize, ops)
    finally:
        lib.cvgs_b200_set_kernel_variant(prev)
    util.assert_bit_equal(general, orc, f"general kernel, seed {seed}")


TYPED = [_abi.CVGS_8UC4, _abi.CVGS_16UC3, _abi.CVGS_16SC4]


@pytest.mark.parametrize("warp_type", [cvgs.WARP_AFFINE, cvgs.WARP_PERSPECTIVE])
@pytest.mark.parametrize("src_type", TYPED)
def test_other_pixel_types_match_reference_kernel_and_oracle(src_type, warp_type):
    """cvGS::warp<WT, InputType> for the other pixel types of the chain (reference include/cvGPUSpeedup.cuh:285-307)."""
    if gpu_util.fkref_lib(16) is None:
        pytest.skip("oracle/_ref not built")
    rng = np.random.default_rng(40 + src_type + warp_type)
    w, h = 211, 157
    px, nc = util.px_bytes_of(src_type), util.channels_of(src_type)
    pitch = px * w + 24
    img = rng.integers(0, 256, size=(h, pitch), dtype=np.uint8)
    d = gpu_util.device_image(img)
    mul = (0.5, 1.25, 1 / 255.0, 2.0)[:nc]
    for m in _matrices(rng, 3, warp_type, w, h):
        inv = cvgs.api.invert_warp_matrix(m, warp_type)
        dsize = (int(rng.integers(8, 300)), int(rng.integers(8, 200)))
        ref = gpu_util.run_fkref_warp_typed(img, w, h, src_type, warp_type, inv, dsize, mul, d_image=d)
        ours = _launch([d], [(w, h)], pitch, [inv], warp_type, dsize, [("mul", mul)], src_type=src_type)[0]
        orc = _oracle([img], [(w, h)], pitch, [inv], warp_type, dsize, [("mul", mul)], src_type=src_type)[0]
        util.assert_bit_equal(ours, ref, f"src {src_type} type {warp_type} dsize {dsize}: ours vs reference kernel")
        util.assert_bit_equal(orc, ref, f"src {src_type} type {warp_type} dsize {dsize}: oracle vs reference kernel")


@pytest.mark.parametrize("src_type", [_abi.CVGS_16SC3, _abi.CVGS_16UC4, _abi.CVGS_8UC4])
def test_other_pixel_types_batches_and_layouts(src_type):
    rng = np.random.default_rng(60 + src_type)
    px, nc = util.px_bytes_of(src_type), util.channels_of(src_type)
    pitch = 1024
    sizes = [(int(rng.integers(8, 120)), int(rng.integers(8, 100))) for _ in range(50)]
    imgs = [rng.integers(0, 256, size=(hh, pitch), dtype=np.uint8) for (_, hh) in sizes]
    d = [gpu_util.device_image(im) for im in imgs]
    inverses = [cvgs.api.invert_warp_matrix(m, cvgs.WARP_PERSPECTIVE) for (w, h) in sizes for m in _matrices(rng, 1, cvgs.WARP_PERSPECTIVE, w, h)]
    ops = [("mul", (0.3, 0.5, 2.0, 1.5)[:nc]), ("sub", (1.0, 4.0, 3.2, 0.5)[:nc]), ("div", (3.2, 0.6, 11.8, 2.0)[:nc])]
    for layout in (_abi.OUT_NCHW, _abi.OUT_NHWC, _abi.OUT_CNHW):
        kw = dict(n_planes=53, used=50, background=(3.0, 5.0, 7.0, 9.0)[:nc], layout=layout, src_type=src_type)
        ours = _launch(d, sizes, pitch, inverses, cvgs.WARP_PERSPECTIVE, (45, 31), ops, **kw)
        orc = _oracle(imgs, sizes, pitch, inverses, cvgs.WARP_PERSPECTIVE, (45, 31), ops, **kw)
        util.assert_bit_equal(ours, orc, f"src {src_type} layout {layout}")


def test_no_used_planes_and_empty_chain():
    """used = 0: every plane is the chain of the default value; nothing is read (images may be NULL)."""
    lib = _abi.load()
    out = torch.full((3, 3, 9, 11), float("nan"), device="cuda")
    p = util.make_pipeline((11, 9), [("mul", (2.0, 3.0, 4.0))], out_ptr=out.data_ptr(), background=(1.0, 2.0, 3.0))
    _abi.check(lib.cvgs_b200_warp_launch(None, None, 3, 0, C.byref(p), None))
    torch.cuda.synchronize()
    got = out.cpu().numpy()
    assert (got[:, 0] == 2.0).all() and (got[:, 1] == 6.0).all() and (got[:, 2] == 12.0).all()


@pytest.mark.parametrize("src_type", [_abi.CVGS_8UC3, _abi.CVGS_8UC4])
def test_warp_into_per_plane_images(src_type):
    """fk::SplitWrite behind a warp: one pitched float image per (plane, channel); equal to the NCHW tensor of the same
    launch, reorder included."""
    rng = np.random.default_rng(90 + src_type)
    px, nc = util.px_bytes_of(src_type), util.channels_of(src_type)
    w, h, pitch = 120, 90, 512
    n = 60  # more than one launch's descriptor table
    img = rng.integers(0, 256, size=(h, pitch), dtype=np.uint8)
    d = gpu_util.device_image(img)
    inverses = [cvgs.api.invert_warp_matrix(m, cvgs.WARP_AFFINE) for m in _matrices(rng, n, cvgs.WARP_AFFINE, w, h)]
    perm = (2, 1, 0) if nc == 3 else (2, 1, 0, 3)
    ops = [("reorder", perm), ("mul", (0.5, 1.5, 2.5, 3.5)[:nc])]
    W, H = 50, 40
    want = _launch([d] * n, [(w, h)] * n, pitch, inverses, cvgs.WARP_AFFINE, (W, H), ops, src_type=src_type)
    lib = _abi.load()
    planes = [torch.full((H, W + (i % 3) * 4), float("nan"), device="cuda") for i in range(n * nc)]
    arr = (_abi.Plane * (n * nc))()
    for i, t in enumerate(planes):
        arr[i].data, arr[i].pitch_bytes = t.data_ptr(), t.stride(0) * 4
    crops = (_abi.Crop * n)()
    warps = (_abi.Warp * n)()
    for i in range(n):
        crops[i].data, crops[i].width, crops[i].height, crops[i].pitch = d.data_ptr(), w, h, pitch
        warps[i].type = cvgs.WARP_AFFINE
        for k in range(9):
            warps[i].m[k] = float(inverses[i][k])
    p = util.make_pipeline((W, H), ops, out_ptr=C.addressof(arr), layout=_abi.OUT_PLANES, src_type=src_type)
    _abi.check(lib.cvgs_b200_warp_launch(crops, warps, n, n, C.byref(p), None))
    torch.cuda.synchronize()
    for z in range(n):
        for c in range(nc):
            util.assert_bit_equal(planes[z * nc + c][:, :W].cpu().numpy(), want[z, c], f"plane {z} channel {c}")


@pytest.mark.parametrize("src_type", [_abi.CVGS_8UC3, _abi.CVGS_8UC4, _abi.CVGS_16UC3, _abi.CVGS_16UC4, _abi.CVGS_16SC3, _abi.CVGS_16SC4])
def test_fast_and_general_warp_kernels_agree(src_type):
    """Chains of the shape [MUL | FMA | ADD] [DIV] with a float tensor behind them take the instantiation without the
    chain interpreter (lane-strided pixels); everything else, and kernel variant 1, the general kernel.  Both against the
    oracle: ragged widths (tail groups of a 128-pixel span), unused planes, every chain shape, strided planes, pixels
    outside the source."""
    nc = util.channels_of(src_type)
    px = util.px_bytes_of(src_type)
    rng = np.random.default_rng(4100 + src_type)
    w, h = 150, 90
    pitch = px * w + 10 * (px // nc)
    img = rng.integers(0, 256, size=(h, pitch), dtype=np.uint8)
    d = gpu_util.device_image(img)
    n = 6
    lib = _abi.load()
    chains = [[], [("mul", (0.5, 0.25, 2.0, 1.5)[:nc])], [("add", (1.0, -2.0, 3.0, 0.5)[:nc])], [("div", (3.2, 0.6, 11.8, 2.0)[:nc])],
              [("mul", (0.3,) * nc), ("sub", (1.0, 4.0, 3.2, 0.5)[:nc]), ("div", (3.2, 0.6, 11.8, 2.0)[:nc])],
              [("reorder", (2, 1, 0, 3)[:nc]), ("sub", (1.0, 4.0, 3.2, 0.5)[:nc]), ("div", (3.0, 7.0, 0.1, 2.5)[:nc])],
              [("div", (3.2, 0.6, 11.8, 2.0)[:nc]), ("mul", (0.3,) * nc)]]  # not the canonical shape: general kernel
    for k, ops in enumerate(chains):
        for warp_type in (cvgs.WARP_AFFINE, cvgs.WARP_PERSPECTIVE):
            dsize = [(131, 37), (128, 8), (33, 9), (257, 3)][k % 4]
            inverses = [cvgs.api.invert_warp_matrix(m, warp_type) for m in _matrices(rng, n, warp_type, w, h)]
            inverses[1] = np.array([1.5, 0.2, -40.0, -0.1, 1.2, -20.0, 0.0, 0.0, 1.0], dtype=np.float32)  # partly outside
            kw = dict(n_planes=n + 2, used=n, background=(7.0, 3.5, 250.0, 1.0)[:nc], src_type=src_type)
            orc = _oracle([img] * n, [(w, h)] * n, pitch, inverses, warp_type, dsize, ops, **kw)
            for variant in (0, 1):
                prev = lib.cvgs_b200_set_kernel_variant(variant)
                try:
                    got = _launch([d] * n, [(w, h)] * n, pitch, inverses, warp_type, dsize, ops, **kw)
                finally:
                    lib.cvgs_b200_set_kernel_variant(prev)
                util.assert_bit_equal(got, orc, f"src {src_type} chain {k} type {warp_type} variant {variant}")
