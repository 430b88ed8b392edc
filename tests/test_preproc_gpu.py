"""GPU parity tests: the sm_100a kernels (through the C-ABI) against the CPU oracle, bit for bit.

Float results are compared bit-exactly (0 ULP): BASELINE.json asks for <= 1 ULP against the
OpenCV-CUDA chain; the kernels reproduce the stated rounding sequence exactly, so the tests demand 0.
"""
import ctypes as C

import numpy as np
import pytest
import torch

from cvgpuspeedup_b200 import _abi
from tests import gpu_util, util

pytestmark = pytest.mark.gpu

VARIANTS = [1, 0]  # 1 = direct-gather kernel, 0 = automatic (TMA-staged when the input allows)


def _check(w: util.Workload, variant, n_planes=None, used=None, **kw):
    got = gpu_util.run_cvgs(w.image, w.rects, w.dsize, w.ops, n_planes=n_planes, used=used, variant=variant,
                            aspect=w.aspect, background=w.background, **kw)
    want = util.run_oracle(w.image, w.rects, w.dsize, w.ops, n_planes=n_planes, used=used, aspect=w.aspect,
                           background=w.background, **kw)
    util.assert_bit_equal(got, want, f"{w.name} variant={variant} {kw}")


@pytest.mark.parametrize("variant", VARIANTS)
@pytest.mark.parametrize("smooth", [False, True])
def test_c1_single_crop(variant, smooth):
    _check(util.workload_c1(smooth=smooth), variant)


@pytest.mark.parametrize("variant", VARIANTS)
@pytest.mark.parametrize("pitch", [6144, 5760])
@pytest.mark.parametrize("ref_shape", [False, True])
def test_c2_fifty_crops(variant, pitch, ref_shape):
    _check(util.workload_c2(pitch=pitch, ref_shape=ref_shape), variant)


@pytest.mark.parametrize("variant", VARIANTS)
def test_c3_imagenet_batch(variant):
    # 96 of the 256 crops keep the oracle's runtime at a few seconds; the full batch is covered by
    # test_full_size_properties
    _check(util.workload_c3(n=96), variant)


@pytest.mark.parametrize("variant", VARIANTS)
@pytest.mark.parametrize("fp", [_abi.FP_REFERENCE_FUSED, _abi.FP_SEPARATE])
@pytest.mark.parametrize("interp", [_abi.INTERP_FLOAT, _abi.INTERP_ROUND_U8])
def test_fp_and_interp_modes(variant, fp, interp):
    _check(util.workload_c2(n=12, frame=(640, 480), pitch=2048), variant, fp_contract=fp, interp_mode=interp)


@pytest.mark.parametrize("variant", VARIANTS)
@pytest.mark.parametrize("aspect", [_abi.PRESERVE_AR, _abi.PRESERVE_AR_RN_EVEN, _abi.PRESERVE_AR_LEFT])
def test_aspect_ratio_modes(variant, aspect):
    rng = np.random.default_rng(5)
    img = util.make_image(rng, 400, 300, pitch=1280)
    rects = [(i, i, 30, 120) for i in range(8)] + [(0, 0, 400, 300), (10, 10, 17, 200), (3, 3, 300, 9)]
    w = util.Workload("ar", img, 400, 300, rects, (64, 128), util.OPS_C2, aspect=aspect,
                      background=(128.0, 128.0, 128.0))
    _check(w, variant)


@pytest.mark.parametrize("variant", VARIANTS)
@pytest.mark.parametrize("layout", [_abi.OUT_NCHW, _abi.OUT_CNHW, _abi.OUT_NHWC])
def test_output_layouts(variant, layout):
    _check(util.workload_c2(n=9, frame=(640, 480), pitch=1920), variant, layout=layout)


@pytest.mark.parametrize("variant", VARIANTS)
def test_padded_plane_stride(variant):
    """explicit batch stride (the reference silently assumes tight)."""
    w = util.workload_c2(n=5, frame=(320, 240), pitch=960)
    _check(w, variant, plane_stride=3 * 64 * 128 + 20)
    _check(w, variant, plane_stride=64 * 128 + 8, layout=_abi.OUT_CNHW)


@pytest.mark.parametrize("variant", VARIANTS)
def test_unused_planes_get_chain_of_background(variant):
    """planes z >= usedPlanes are written with chain(background), not skipped."""
    w = util.workload_c2(n=6, frame=(320, 240), pitch=960)
    w.background = (3.0, 200.0, 77.5)
    _check(w, variant, n_planes=10, used=4)
    _check(w, variant, n_planes=3, used=0)


@pytest.mark.parametrize("variant", VARIANTS)
@pytest.mark.parametrize("dsize", [(63, 17), (1, 1), (5, 300), (130, 2), (224, 224), (66, 66)])
def test_ragged_destination_sizes(variant, dsize):
    rng = np.random.default_rng(7)
    img = util.make_image(rng, 333, 211, pitch=1003)  # odd pitch: rows at every byte alignment
    rects = [(0, 0, 333, 211), (1, 1, 1, 1), (332, 210, 1, 1), (0, 0, 2, 2), (7, 9, 100, 3), (300, 5, 33, 200),
             (10, 10, 64, 128), (11, 12, 65, 129)]
    w = util.Workload("ragged", img, 333, 211, rects, dsize, util.OPS_C1)
    _check(w, variant)


@pytest.mark.parametrize("variant", VARIANTS)
def test_right_and_bottom_edge_clamp(variant):
    """x2_read/y2_read clamping (interpolation.cuh:72-74) on crops that end at the last byte of the
    allocation: up-scaling makes the last columns/rows tap beyond the crop."""
    rng = np.random.default_rng(8)
    img = util.make_image(rng, 48, 40)
    rects = [(0, 0, 48, 40), (40, 30, 8, 10), (47, 39, 1, 1), (45, 0, 3, 40), (0, 37, 48, 3)]
    w = util.Workload("edge", img, 48, 40, rects, (96, 100), util.OPS_C1)
    _check(w, variant)


@pytest.mark.parametrize("variant", VARIANTS)
def test_chain_shapes(variant):
    w = util.workload_c2(n=4, frame=(320, 240), pitch=960)
    chains = [[], [("div", (255.0,) * 3)], [("add", (1.5, 2.5, 3.5)), ("mul", (2.0, 3.0, 4.0))],
              [("mul", (0.5,) * 3), ("reorder", (2, 1, 0)), ("sub", (1.0, 2.0, 3.0))],
              [("reorder", (1, 2, 0)), ("mul", (0.25, 0.5, 2.0)), ("add", (1.0, 2.0, 3.0)), ("add", (0.1, 0.2, 0.3)),
               ("div", (3.0, 7.0, 0.1)), ("reorder", (2, 0, 1)), ("sub", (5.0, 6.0, 7.0)), ("mul", (1.1, 1.2, 1.3))]]
    for ops in chains:
        w.ops = ops
        _check(w, variant)
        _check(w, variant, fp_contract=_abi.FP_SEPARATE)


def test_large_batch_uses_descriptor_ring():
    """More crops than fit the kernel parameters (64): descriptors go through the pinned staging ring."""
    w = util.workload_c2(n=300, frame=(640, 480), pitch=1920)
    w.dsize = (32, 48)
    for variant in VARIANTS:
        for _ in range(3):  # reuse of ring slots
            _check(w, variant)


def test_full_size_properties():
    """BASELINE configs 2/3 at full size through size-independent properties: (1) both kernels agree bit for
    bit, (2) the result does not depend on how the batch is split into launches (planes are independent),
    (3) linearity of the chain: doubling mul/sub doubles the output exactly (power-of-two scaling)."""
    w = util.workload_c3(n=256)
    d_img = gpu_util.device_image(w.image)
    a = gpu_util.run_cvgs(w.image, w.rects, w.dsize, w.ops, variant=1, d_image=d_img)
    b = gpu_util.run_cvgs(w.image, w.rects, w.dsize, w.ops, variant=0, d_image=d_img)
    util.assert_bit_equal(a, b, "direct vs auto kernel, 256 crops")
    parts = [gpu_util.run_cvgs(w.image, w.rects[i:i + 50], w.dsize, w.ops, d_image=d_img) for i in range(0, 256, 50)]
    util.assert_bit_equal(np.concatenate(parts), b, "batch split invariance")
    ops2 = [("reorder", (2, 1, 0)), ("mul", tuple(2 * v for v in (1 / 255.0,) * 3)),
            ("sub", tuple(2 * v for v in util._MEAN)), ("div", util._STD)]
    c = gpu_util.run_cvgs(w.image, w.rects, w.dsize, ops2, d_image=d_img)
    util.assert_bit_equal(c, 2 * b, "exact scaling by 2")
    # spot-check 8 planes of the full batch against the oracle
    idx = [0, 1, 37, 100, 128, 200, 254, 255]
    want = util.run_oracle(w.image, [w.rects[i] for i in idx], w.dsize, w.ops)
    util.assert_bit_equal(b[idx], want, "oracle spot check")


def test_error_behaviour():
    """Bad arguments come back as error codes + message (the shim rethrows, like gpuErrchk)."""
    lib = _abi.load()
    p = util.make_pipeline((64, 128), [], out_ptr=0)
    crops = (_abi.Crop * 1)()
    assert lib.cvgs_b200_preproc_launch(crops, 1, 1, C.byref(p), None) == 1
    assert b"output" in lib.cvgs_b200_last_error()
    d_out = torch.zeros(3 * 64 * 128, device="cuda")
    p = util.make_pipeline((64, 128), [], out_ptr=d_out.data_ptr())
    assert lib.cvgs_b200_preproc_launch(crops, 1, 1, C.byref(p), None) == 1  # crop.data == NULL
    assert b"crop 0" in lib.cvgs_b200_last_error()
    p.src_type = 0
    assert lib.cvgs_b200_preproc_launch(crops, 1, 1, C.byref(p), None) == 801


def test_host_buffer_entry_point():
    """cvgs_b200_preproc_host: pinned host frame in, host tensor out, copies on the caller's stream."""
    lib = _abi.load()
    w = util.workload_c2(n=20, frame=(640, 480), pitch=1920)
    h_img = torch.from_numpy(w.image).pin_memory()
    h_out = torch.empty((20, 3, 128, 64), dtype=torch.float32).pin_memory()
    rects = (_abi.Rect * 20)(*[_abi.Rect(*r) for r in w.rects])
    p = util.make_pipeline(w.dsize, w.ops)
    for _ in range(2):
        h_out.fill_(float("nan"))
        _abi.check(lib.cvgs_b200_preproc_host(h_img.data_ptr(), 640, 480, 1920, rects, 20, 20, C.byref(p),
                                              h_out.data_ptr(), torch.cuda.current_stream().cuda_stream))
        torch.cuda.synchronize()
        util.assert_bit_equal(h_out.numpy(), util.run_oracle(w.image, w.rects, w.dsize, w.ops), "host entry point")


@pytest.mark.parametrize("src_type,px", [(_abi.CVGS_8UC3, 3), (_abi.CVGS_16UC3, 6), (_abi.CVGS_8UC4, 4)])
def test_rectangles_of_a_device_frame(src_type, px):
    """cvgs_b200_preproc_launch_rects (cvGS::crop(read, rects)[.then(resize)], fk::Crop crop.cuh:23-55): a device frame +
    a rectangle list equals the ROI-pointer launch; partial batches, and rectangles outside the frame are refused."""
    lib = _abi.load()
    rng = np.random.default_rng(91)
    fw, fh = 400, 300
    pitch = (px * fw + 63) // 64 * 64
    img = rng.integers(0, 256, size=(fh, pitch), dtype=np.uint8)
    rects = [(0, 0, 400, 300), (399, 299, 1, 1), (17, 23, 100, 50), (200, 10, 64, 128), (3, 200, 333, 77), (50, 60, 70, 80)]
    n = len(rects) + 2
    nc = util.channels_of(src_type)
    ops = [("mul", (0.5,) * nc), ("sub", (1.0, 2.0, 3.0, 4.0)[:nc])]
    d_img = torch.from_numpy(img).cuda()
    out = torch.full((n, nc, 128, 64), float("nan"), device="cuda")
    p = util.make_pipeline((64, 128), ops, out_ptr=out.data_ptr(), src_type=src_type, background=(9, 8, 7, 6)[:nc])
    r_arr = (_abi.Rect * len(rects))(*[_abi.Rect(*r) for r in rects])
    _abi.check(lib.cvgs_b200_preproc_launch_rects(d_img.data_ptr(), fw, fh, pitch, r_arr, n, len(rects), C.byref(p), None))
    torch.cuda.synchronize()
    want = util.run_oracle(img, rects, (64, 128), ops, n_planes=n, used=len(rects), src_type=src_type, background=(9, 8, 7, 6)[:nc])
    util.assert_bit_equal(out.cpu().numpy(), want, "rectangles of a device frame")
    bad = (_abi.Rect * 1)(_abi.Rect(390, 0, 20, 20))
    assert lib.cvgs_b200_preproc_launch_rects(d_img.data_ptr(), fw, fh, pitch, bad, 1, 1, C.byref(p), None) == 1
    assert b"outside" in lib.cvgs_b200_last_error()


@pytest.fixture()
def tile_upload():
    lib = _abi.load()
    prev = lib.cvgs_b200_set_host_upload(1)
    yield lib
    lib.cvgs_b200_set_host_upload(prev)


@pytest.mark.parametrize("pinned", [True, False])
@pytest.mark.parametrize("frame", [(640, 480, 1920), (333, 211, 1008), (1000, 37, 3008)])
def test_host_buffer_upload_paths(tile_upload, pinned, frame):
    """cvgs_b200_set_host_upload(1): pinned frames are pulled tile by tile by the upload kernel (frame widths that are /
    are not multiples of the 128-byte tile, heights that are not multiples of 16 rows); pageable frames take the row
    copy.  Same results."""
    lib = tile_upload
    fw, fh, pitch = frame
    w = util.workload_c2(seed=77, n=12, frame=(fw, fh), pitch=pitch)
    w.rects = [(x, y, max(1, min(ww, fw - x)), max(1, min(hh, fh - y))) for (x, y, ww, hh) in w.rects] + [(0, 0, fw, fh), (fw - 1, fh - 1, 1, 1)]
    n = len(w.rects)
    h_img = torch.from_numpy(w.image.copy())
    if pinned:
        h_img = h_img.pin_memory()
    h_out = torch.empty((n, 3, 128, 64), dtype=torch.float32).pin_memory()
    rects = (_abi.Rect * n)(*[_abi.Rect(*r) for r in w.rects])
    p = util.make_pipeline(w.dsize, w.ops)
    lib.cvgs_b200_debug_host_bytes(None, None, 1)
    h_out.fill_(float("nan"))
    _abi.check(lib.cvgs_b200_preproc_host(h_img.data_ptr(), fw, fh, pitch, rects, n, n, C.byref(p), h_out.data_ptr(),
                                          torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    up, down = C.c_uint64(), C.c_uint64()
    lib.cvgs_b200_debug_host_bytes(C.byref(up), C.byref(down), 1)
    assert down.value == h_out.numel() * 4 and up.value >= 3 * fw * fh
    util.assert_bit_equal(h_out.numpy(), util.run_oracle(w.image, w.rects, w.dsize, w.ops), "host entry point")


@pytest.mark.parametrize("ops,nc", [([("reorder", (2, 1, 0)), ("add_alpha", (255.0,)), ("mul", (0.5, 0.25, 2.0, 1.0))], 4),
                                    ([("gray", (1,)), ("mul", (0.5,))], 1)])
def test_host_buffer_entry_point_with_channel_changing_chain(ops, nc):
    """cvtColor<BGR2RGBA> / <RGB2GRAY> in the host-buffer path: the tensor has the channels the chain leaves (4 / 1)."""
    lib = _abi.load()
    w = util.workload_c2(seed=78, n=9, frame=(320, 240), pitch=960)
    h_img = torch.from_numpy(w.image).pin_memory()
    guard = 1024
    h_out = torch.full((9 * nc * 128 * 64 + guard,), -7.0, dtype=torch.float32).pin_memory()
    rects = (_abi.Rect * 9)(*[_abi.Rect(*r) for r in w.rects])
    p = util.make_pipeline(w.dsize, ops)
    _abi.check(lib.cvgs_b200_preproc_host(h_img.data_ptr(), 320, 240, 960, rects, 9, 9, C.byref(p), h_out.data_ptr(),
                                          torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    got = h_out.numpy()
    assert np.all(got[9 * nc * 128 * 64:] == -7.0), "the download ran past the tensor"
    util.assert_bit_equal(got[:9 * nc * 128 * 64].reshape(9, nc, 128, 64), util.run_oracle(w.image, w.rects, w.dsize, ops),
                          "host entry point, channel-changing chain")


# ---- the TMA-staged kernel, forced (variant 2 fails loudly instead of falling back) ----
def test_tma_kernel_c1_c2_c3():
    _check(util.workload_c1(), 2)                      # 10x / 3.75x down-scale
    _check(util.workload_c1(smooth=True), 2)
    _check(util.workload_c2(pitch=6144), 2)            # mixed up- and down-scales, 50 crops
    _check(util.workload_c2(pitch=5760, ref_shape=True), 2)
    w = util.workload_c3(n=40)                         # 224x224: two column bands per crop
    _check(w, 2)


@pytest.mark.parametrize("dsize", [(63, 17), (1, 1), (5, 300), (130, 2), (224, 224), (66, 66), (300, 40), (640, 9)])
def test_tma_kernel_ragged_sizes(dsize):
    rng = np.random.default_rng(17)
    img = util.make_image(rng, 333, 211, pitch=1008)
    rects = [(0, 0, 333, 211), (1, 1, 1, 1), (332, 210, 1, 1), (0, 0, 2, 2), (7, 9, 100, 3), (300, 5, 33, 200),
             (10, 10, 64, 128), (11, 12, 65, 129), (5, 0, 328, 1)]
    w = util.Workload("ragged_tma", img, 333, 211, rects, dsize, util.OPS_C1)
    _check(w, 2)
    _check(w, 2, fp_contract=_abi.FP_SEPARATE, interp_mode=_abi.INTERP_ROUND_U8)


@pytest.mark.parametrize("aspect", [_abi.PRESERVE_AR, _abi.PRESERVE_AR_RN_EVEN, _abi.PRESERVE_AR_LEFT])
def test_tma_kernel_aspect_modes_and_layouts(aspect):
    rng = np.random.default_rng(5)
    img = util.make_image(rng, 400, 300, pitch=1280)
    rects = [(i, i, 30, 120) for i in range(8)] + [(0, 0, 400, 300), (10, 10, 17, 200), (3, 3, 300, 9)]
    for dsize in [(64, 128), (300, 100)]:
        w = util.Workload("ar_tma", img, 400, 300, rects, dsize, util.OPS_C2, aspect=aspect,
                          background=(128.0, 7.0, 250.5))
        _check(w, 2)
        _check(w, 2, layout=_abi.OUT_NHWC, n_planes=14, used=9)
        _check(w, 2, layout=_abi.OUT_CNHW, plane_stride=dsize[0] * dsize[1] + 4)


def test_tma_kernel_unaligned_base_and_chains():
    """ROI pointers at every 16-byte phase (image base shifted by 0..15 bytes) and every chain shape."""
    rng = np.random.default_rng(23)
    back = rng.integers(0, 256, size=240 * 960 + 64, dtype=np.uint8)
    d_back = torch.from_numpy(back).cuda()
    rects = [(0, 0, 319, 240), (3, 5, 40, 80), (100, 20, 200, 200), (318, 0, 1, 240), (7, 7, 64, 128)]
    for shift in [0, 1, 5, 8, 15]:
        img = back[shift:shift + 240 * 960].reshape(240, 960)
        d_img = d_back[shift:shift + 240 * 960].view(240, 960)
        w = util.Workload("shift", img, 319, 240, rects, (64, 128), util.OPS_C2)
        got = gpu_util.run_cvgs(img, rects, w.dsize, w.ops, variant=2, d_image=d_img)
        util.assert_bit_equal(got, util.run_oracle(img, rects, w.dsize, w.ops), f"base shift {shift}")
    w = util.workload_c2(n=4, frame=(320, 240), pitch=960)
    chains = [[], [("div", (255.0,) * 3)], [("add", (1.5, 2.5, 3.5)), ("mul", (2.0, 3.0, 4.0))],
              [("mul", (0.5,) * 3), ("reorder", (2, 1, 0)), ("sub", (1.0, 2.0, 3.0))],
              [("mul", (1e-30,) * 3), ("mul", (1e30,) * 3)],
              [("reorder", (1, 2, 0)), ("mul", (0.25, 0.5, 2.0)), ("add", (1.0, 2.0, 3.0)), ("add", (0.1, 0.2, 0.3)),
               ("div", (3.0, 7.0, 0.1)), ("reorder", (2, 0, 1)), ("sub", (5.0, 6.0, 7.0)), ("mul", (1.1, 1.2, 1.3))]]
    for ops in chains:
        w.ops = ops
        _check(w, 2)
        _check(w, 2, fp_contract=_abi.FP_SEPARATE)
        _check(w, 2, interp_mode=_abi.INTERP_ROUND_U8)


def test_tma_kernel_large_batch_device_tables():
    """More than 64 crops: tensor maps and descriptors travel through the pinned ring into device memory."""
    w = util.workload_c2(n=300, frame=(640, 480), pitch=1920)
    w.dsize = (32, 48)
    for _ in range(10):  # more launches than ring slots: slot (and tensor-map address) reuse
        _check(w, 2)
    w2 = util.workload_c2(seed=9, n=300, frame=(640, 480), pitch=1920)
    w2.dsize = (32, 48)
    _check(w2, 2)


def test_tma_kernel_refuses_unsupported_pitch():
    w = util.workload_c2(n=3, frame=(333, 211), pitch=1003)
    with pytest.raises(_abi.CvgsError):
        _check(w, 2)
    _check(w, 0)  # automatic selection falls back to the direct-gather kernel


@pytest.mark.parametrize("variant", [1, 2])
def test_identity_scale_is_a_plain_batch_read(variant):
    """Destination size == source size (fk::BatchRead<PerThreadRead>, the reference's batch reads without a resize,
    tests/batchread/test_batchread_x_write3D.cu): scale factors are exactly 1, so the output is the source pixel."""
    rng = np.random.default_rng(77)
    img = util.make_image(rng, 200, 120, pitch=640)
    rects = [(3 * i, 2 * i, 96, 64) for i in range(10)]
    got = gpu_util.run_cvgs(img, rects, (96, 64), [], variant=variant, layout=_abi.OUT_NHWC)
    px = img[:, :600].reshape(120, 200, 3).astype(np.float32)
    for i, (x, y, w, h) in enumerate(rects):
        util.assert_bit_equal(got[i], px[y:y + h, x:x + w], f"crop {i}")
