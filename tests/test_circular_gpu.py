"""CircularTensor update kernel vs the oracle state machine and the reference's known answers."""
import ctypes as C

import numpy as np
import pytest
import torch

from cvgpuspeedup_b200 import _abi
import cvgpuspeedup_b200 as cvgs
from tests import util

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("order,expect", [(_abi.CT_NEWEST_FIRST, lambda z: 100 - z),
                                          (_abi.CT_OLDEST_FIRST, lambda z: 100 - (15 - z - 1))])
@pytest.mark.parametrize("mode", [_abi.CT_STANDARD, _abi.CT_TRANSPOSED])
def test_known_answer_after_100_updates(order, expect, mode):
    """tests/batchread/test_circularbatchread_x_write3D.cu:176-270,333-460: 128x128, BATCH 15, 100 updates with
    frame value i+1; plane z == 100-z (NewestFirst) / 100-(15-z-1) (OldestFirst)."""
    ct = cvgs.CircularTensor(128, 128, 15, order, mode)
    frame = torch.empty((128, 128, 3), dtype=torch.uint8, device="cuda")
    for i in range(100):
        frame.fill_(i + 1)
        ct.update(None, cvgs.GpuMat.from_tensor(frame))
    torch.cuda.synchronize()
    data = ct.data().cpu().numpy()
    if mode == _abi.CT_TRANSPOSED:
        data = data.swapaxes(0, 1)
    for z in range(15):
        assert np.all(data[z] == expect(z)), z
    ct.close()


@pytest.mark.parametrize("order", [_abi.CT_NEWEST_FIRST, _abi.CT_OLDEST_FIRST])
@pytest.mark.parametrize("mode", [_abi.CT_STANDARD, _abi.CT_TRANSPOSED])
@pytest.mark.parametrize("shape", [((96, 64), (200, 120)), ((33, 17), (33, 17)), ((64, 36), (192, 108))])
def test_matches_oracle_with_resize_chain(order, mode, shape):
    """BASELINE config 4 in miniature: random frames, resize + normalise on the new frame, depth 5."""
    (W, H), (fw, fh) = shape
    B = 5
    lib = util.oracle_lib()
    o = lib.oracle_ct_create(W, H, 3, B, order, mode)
    ct = cvgs.CircularTensor(W, H, B, order, mode)
    rng = np.random.default_rng(4)
    p = util.make_pipeline((W, H), util.OPS_C3)
    ops = [cvgs.cvtColor(), cvgs.multiply((1 / 255.0,) * 3), cvgs.subtract(util._MEAN), cvgs.divide(util._STD)]
    for i in range(2 * B + 3):
        img = util.make_image(rng, fw, fh)
        assert lib.oracle_ct_update(o, util.host_crops(img, [(0, 0, fw, fh)]), C.byref(p), 0) == 0
        d = torch.from_numpy(img).cuda().view(fh, fw, 3)
        ct.update(None, cvgs.GpuMat.from_tensor(d), *ops)
        torch.cuda.synchronize()
        want = np.ctypeslib.as_array(lib.oracle_ct_data(o), shape=(B * 3 * H * W,))
        got = ct.data().cpu().numpy().reshape(-1)
        util.assert_bit_equal(got, want, f"update {i}")
    lib.oracle_ct_destroy(o)
    ct.close()


def test_errors():
    lib = _abi.load()
    h = C.c_void_p()
    assert lib.cvgs_b200_ct_create(C.byref(h), 0, 10, 3, 4, 0, 0, 0) == 1
    assert lib.cvgs_b200_ct_create(C.byref(h), 16, 16, 2, 4, 0, 0, 0) == 801  # 1, 3 or 4 colour planes
    ct = cvgs.CircularTensor(16, 16, 3)
    frame = torch.zeros((16, 16, 3), dtype=torch.uint8, device="cuda")
    p = util.make_pipeline((8, 8), [])
    crop = util.host_crops(np.zeros((16, 48), np.uint8), [(0, 0, 16, 16)], base_ptr=frame.data_ptr())
    assert lib.cvgs_b200_ct_update(ct._h, crop, C.byref(p), None) == 1  # wrong destination size
    ct.close()


@pytest.mark.parametrize("seed", range(10))
def test_random_depths_sizes_and_orders(seed):
    """Depth 1 (nothing to shift) to 9, plane sizes with and without 16-byte multiples, every order / layout, a run of
    updates longer than the ring."""
    rng = np.random.default_rng(700 + seed)
    W, H = int(rng.integers(1, 90)), int(rng.integers(1, 60))
    fw, fh = int(rng.integers(2, 200)), int(rng.integers(2, 150))
    B = int(rng.integers(1, 10))
    order = int(rng.integers(0, 2))
    mode = int(rng.integers(0, 2))
    lib = util.oracle_lib()
    o = lib.oracle_ct_create(W, H, 3, B, order, mode)
    ct = cvgs.CircularTensor(W, H, B, order, mode)
    p = util.make_pipeline((W, H), util.OPS_C2)
    ops = [cvgs.cvtColor(), cvgs.multiply((0.3, 0.3, 0.3)), cvgs.subtract((1.0, 4.0, 3.2)), cvgs.divide((3.2, 0.6, 11.8))]
    for i in range(B + 4):
        img = util.make_image(rng, fw, fh)
        assert lib.oracle_ct_update(o, util.host_crops(img, [(0, 0, fw, fh)]), C.byref(p), 0) == 0
        d = torch.from_numpy(img).cuda().view(fh, fw, 3)
        ct.update(None, cvgs.GpuMat.from_tensor(d), *ops)
    torch.cuda.synchronize()
    want = np.ctypeslib.as_array(lib.oracle_ct_data(o), shape=(B * 3 * H * W,))
    util.assert_bit_equal(ct.data().cpu().numpy().reshape(-1), want, f"seed {seed}: {W}x{H} depth {B} order {order} mode {mode}")
    lib.oracle_ct_destroy(o)
    ct.close()


def _fkref_ct():
    import os
    path = os.path.join(util.ROOT, "oracle", "_ref", "libfkref_ct.so")
    if not os.path.exists(path):
        pytest.skip("oracle/_ref/libfkref_ct.so not built (needs /root/reference at build time)")
    ref = C.CDLL(path)
    ref.fkref_ct_create.restype = C.c_void_p
    ref.fkref_ct_create.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int]
    ref.fkref_ct_update.restype = C.c_int
    ref.fkref_ct_update.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_float),
                                    C.POINTER(C.c_float), C.POINTER(C.c_float), C.c_void_p]
    ref.fkref_ct_data.restype = C.c_void_p
    ref.fkref_ct_data.argtypes = [C.c_void_p]
    ref.fkref_ct_bytes.restype = C.c_ulonglong
    ref.fkref_ct_bytes.argtypes = [C.c_void_p]
    ref.fkref_ct_destroy.argtypes = [C.c_void_p]
    return ref


@pytest.mark.parametrize("order", [_abi.CT_NEWEST_FIRST, _abi.CT_OLDEST_FIRST])
@pytest.mark.parametrize("depth,shape", [(4, ((96, 64), (200, 120))), (4, ((64, 48), (64, 48))), (16, ((160, 90), (640, 360))),
                                         (16, ((320, 180), (320, 180)))])
@pytest.mark.parametrize("swap", [0, 1])
def test_three_way_against_the_reference_kernel(order, depth, shape, swap):
    """fk::CircularTensor<float, 3, BATCH, ORDER, Standard>::update instantiated from the reference's own headers
    (oracle/_ref/libfkref_ct.so; circular_tensor.cuh:111-146), this library's kernel and the oracle state machine on
    the same frames, after every update once the ring has wrapped.

    Plane-sized frames (Read + SaturateCast, the form of the reference's own CircularTensor tests): all three bit-equal.
    Frames that are resized: ours == oracle bit for bit; the reference's instantiation agrees bit for bit in the
    channels where nvcc contracted p00*w00 + p10*w10 + p01*w01 + p11*w11 the way it does in the batch kernels
    (FMUL(p10*w10) first), and differs by one rounding of the interpolated value in the channel(s) where this particular
    instantiation starts with FMUL(p00*w00) instead (SASS of libfkref_ct.so) -- at most 1e-6 after the chain."""
    ref = _fkref_ct()
    (W, H), (fw, fh) = shape
    resized = (W, H) != (fw, fh)
    h = ref.fkref_ct_create(depth, order, W, H)
    assert h
    assert ref.fkref_ct_bytes(h) == 4 * 3 * depth * W * H
    lib = util.oracle_lib()
    o = lib.oracle_ct_create(W, H, 3, depth, order, _abi.CT_STANDARD)
    ct = cvgs.CircularTensor(W, H, depth, order, _abi.CT_STANDARD)
    ref_view = cvgs.api.device_view(ref.fkref_ct_data(h), (depth, 3, H, W))
    rng = np.random.default_rng(40 + depth)
    chain = util.OPS_C3 if swap else util.OPS_C3[1:]
    p = util.make_pipeline((W, H), chain)
    ops = ([cvgs.cvtColor()] if swap else []) + [cvgs.multiply((1 / 255.0,) * 3), cvgs.subtract(util._MEAN), cvgs.divide(util._STD)]
    f3 = lambda v: (C.c_float * 3)(*v)  # noqa: E731
    mul, sub, div = f3((1 / 255.0,) * 3), f3(util._MEAN), f3(util._STD)
    st = torch.cuda.current_stream().cuda_stream
    for i in range(depth + 3):
        img = util.make_image(rng, fw, fh)
        d = torch.from_numpy(img).cuda().view(fh, fw, 3)
        assert lib.oracle_ct_update(o, util.host_crops(img, [(0, 0, fw, fh)]), C.byref(p), 0) == 0
        ct.update(None, cvgs.GpuMat.from_tensor(d), *ops)
        assert ref.fkref_ct_update(h, d.data_ptr(), fw, fh, 3 * fw, swap, mul, sub, div, st) == 0
        torch.cuda.synchronize()
        if i < depth - 1:  # until the ring has wrapped, the reference's temp tensor holds uninitialised planes
            continue
        want = np.ctypeslib.as_array(lib.oracle_ct_data(o), shape=(depth, 3, H, W))
        util.assert_bit_equal(ct.data().cpu().numpy(), want, f"update {i}: ours vs oracle")
        got_ref = ref_view.cpu().numpy()
        if not resized:
            util.assert_bit_equal(got_ref, want, f"update {i}: reference kernel vs oracle")
            continue
        exact = [bool(np.array_equal(got_ref[:, c].view(np.uint32), want[:, c].view(np.uint32))) for c in range(3)]
        assert sum(exact) >= 1, f"update {i}: no channel of the reference's instantiation follows the batch kernels' order"
        assert float(np.abs(got_ref - want).max()) <= 1e-6, f"update {i}: more than one rounding apart"
    lib.oracle_ct_destroy(o)
    ct.close()
    ref.fkref_ct_destroy(h)


@pytest.mark.parametrize("case", [
    # (frame type, px bytes, colour planes, element channels, chain)  -- reference include/cvGPUSpeedup.cuh:600-627
    ("8UC4->4 planes", _abi.CVGS_8UC4, 4, 4, 1, [("mul", (0.5, 0.25, 2.0, 1.0)), ("sub", (1.0, 2.0, 3.0, 4.0))]),
    ("8UC4->packed float4 (test_circularbatchread_x_write3D.cu:400-460)", _abi.CVGS_8UC4, 4, 1, 4, [("mul", (0.5, 0.25, 2.0, 1.0))]),
    ("8UC3->packed float3", _abi.CVGS_8UC3, 3, 1, 3, [("reorder", (2, 1, 0)), ("mul", (0.5, 0.25, 2.0))]),
    ("8UC3->gray, one plane", _abi.CVGS_8UC3, 3, 1, 1, [("gray", (1,)), ("mul", (0.5,))]),
    ("16UC4->4 planes", _abi.CVGS_16UC4, 8, 4, 1, [("mul", (0.001, 0.002, 0.003, 0.004))]),
    ("8UC3->RGBA planes", _abi.CVGS_8UC3, 3, 4, 1, [("add_alpha", (255.0,)), ("mul", (0.5, 0.5, 0.5, 0.5))]),
], ids=lambda c: c[0])
@pytest.mark.parametrize("order", [_abi.CT_NEWEST_FIRST, _abi.CT_OLDEST_FIRST])
@pytest.mark.parametrize("mode", [_abi.CT_STANDARD, _abi.CT_TRANSPOSED])
def test_other_colour_plane_counts_and_packed_elements(case, order, mode):
    """COLOR_PLANES 1 / 4, packed float3 / float4 elements and 4-channel frames against the oracle state machine."""
    _name, src_type, px, cp, ec, ops = case
    W, H, fw, fh, B = 40, 24, 100, 60, 3
    lib, olib = _abi.load(), util.oracle_lib()
    o = olib.oracle_ct_create_ex(W, H, cp, ec, B, order, mode)
    h = C.c_void_p()
    _abi.check(lib.cvgs_b200_ct_create_ex(C.byref(h), W, H, cp, ec, B, order, mode, 0))
    rng = np.random.default_rng(8)
    p = util.make_pipeline((W, H), ops, src_type=src_type)
    n_floats = B * cp * ec * H * W
    for i in range(2 * B + 1):
        img = rng.integers(0, 256, size=(fh, px * fw), dtype=np.uint8)
        d = torch.from_numpy(img).cuda()
        assert olib.oracle_ct_update(o, util.host_crops(img, [(0, 0, fw, fh)], px_bytes=px), C.byref(p), 0) == 0
        _abi.check(lib.cvgs_b200_ct_update(h, util.host_crops(img, [(0, 0, fw, fh)], base_ptr=d.data_ptr(), px_bytes=px), C.byref(p), None))
        torch.cuda.synchronize()
        got = cvgs.api.device_view(lib.cvgs_b200_ct_data(h), (n_floats,)).cpu().numpy()
        want = np.ctypeslib.as_array(olib.oracle_ct_data(o), shape=(n_floats,))
        util.assert_bit_equal(got, want, f"update {i}")
    # a chain that ends with another channel count is refused
    bad = util.make_pipeline((W, H), [("gray", (1,))] if cp * ec != 1 else [], src_type=src_type)
    crop = util.host_crops(img, [(0, 0, fw, fh)], base_ptr=d.data_ptr(), px_bytes=px)
    assert lib.cvgs_b200_ct_update(h, crop, C.byref(bad), None) == 1 and b"channels" in lib.cvgs_b200_last_error()
    olib.oracle_ct_destroy(o)
    lib.cvgs_b200_ct_destroy(h)
