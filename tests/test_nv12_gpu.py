"""NV12 frames (decoder output) as the source of the fused path: fk::ReadYUV<NV12> + fk::ConvertYUVToRGB in front of
the resize (reference color_conversion.cuh:235-362; tests/resize/test_fused_resize.cu:73-76,141-143 -- a test without
assertions there).  Bit-exact against the reference's own kernel and the oracle, for the four matrices it defines."""
import ctypes as C

import numpy as np
import pytest
import torch

from cvgpuspeedup_b200 import _abi
from tests import gpu_util, util

pytestmark = pytest.mark.gpu

MUL, SUB, DIV = (1 / 255.0,) * 3, (0.485, 0.456, 0.406), (0.229, 0.224, 0.225)
OPS = [("mul", MUL), ("sub", SUB), ("div", DIV)]


def _frame(seed, w, h, pitch):
    rng = np.random.default_rng(seed)
    return rng.integers(0, 256, size=(h + (h + 1) // 2, pitch), dtype=np.uint8)


def _ours(frames, sizes, pitch, dsize, standard, ops, src_type=_abi.CVGS_NV12, **kw):
    """One launch for all frames (each frame is one 'crop' = a whole NV12 image)."""
    lib = _abi.load()
    d = [torch.from_numpy(f).cuda() for f in frames]
    n = len(frames)
    crops = (_abi.Crop * n)()
    for i, (t, (w, h)) in enumerate(zip(d, sizes)):
        crops[i].data, crops[i].width, crops[i].height, crops[i].pitch = t.data_ptr(), w, h, pitch
    out = torch.full((n, 3, dsize[1], dsize[0]), float("nan"), device="cuda")
    p = util.make_pipeline(dsize, ops, out_ptr=out.data_ptr(), src_type=src_type, yuv_standard=standard, **kw)
    _abi.check(lib.cvgs_b200_preproc_launch(crops, n, n, C.byref(p), None))
    torch.cuda.synchronize()
    return out.cpu().numpy()


def _oracle(frames, sizes, pitch, dsize, standard, ops, src_type=_abi.CVGS_NV12, **kw):
    n = len(frames)
    crops = (_abi.Crop * n)()
    for i, (f, (w, h)) in enumerate(zip(frames, sizes)):
        crops[i].data, crops[i].width, crops[i].height, crops[i].pitch = f.ctypes.data, w, h, pitch
    out = np.full((n, 3, dsize[1], dsize[0]), np.nan, dtype=np.float32)
    p = util.make_pipeline(dsize, ops, out_ptr=out.ctypes.data, src_type=src_type, yuv_standard=standard, **kw)
    assert util.oracle_lib().oracle_preproc(crops, n, n, C.byref(p), 0) == 0
    return out


@pytest.mark.parametrize("standard", [0, 1, 2, 3])
def test_nv12_matches_reference_kernel_and_oracle(standard):
    if gpu_util.fkref_lib(16) is None:
        pytest.skip("oracle/_ref not built")
    pitch = 512
    sizes = [(320, 240), (322, 242), (161, 121), (64, 36), (2, 2)]
    frames = [_frame(10 * standard + i, w, h, pitch) for i, (w, h) in enumerate(sizes)]
    for dsize in [(64, 128), (400, 300), (33, 7)]:
        ours = _ours(frames, sizes, pitch, dsize, standard, OPS)
        orc = _oracle(frames, sizes, pitch, dsize, standard, OPS)
        util.assert_bit_equal(ours, orc, f"standard {standard} dsize {dsize}: ours vs oracle")
        for i, (f, (w, h)) in enumerate(zip(frames, sizes)):
            ref = gpu_util.run_fkref_nv12(f, w, h, dsize, standard, MUL, SUB, DIV)
            util.assert_bit_equal(ours[i], ref, f"standard {standard} dsize {dsize} frame {i}: ours vs reference kernel")


def test_nv12_modes_and_layouts():
    pitch = 384
    sizes = [(300, 200), (128, 128)]
    frames = [_frame(77 + i, w, h, pitch) for i, (w, h) in enumerate(sizes)]
    ops = [("reorder", (2, 1, 0)), ("mul", (0.5, 0.25, 2.0)), ("add", (1.0, 2.0, 3.0))]
    for kw in [dict(aspect=_abi.PRESERVE_AR, background=(1.0, 2.0, 3.0)), dict(interp_mode=_abi.INTERP_ROUND_U8),
               dict(fp_contract=_abi.FP_SEPARATE)]:
        util.assert_bit_equal(_ours(frames, sizes, pitch, (96, 64), 1, ops, **kw), _oracle(frames, sizes, pitch, (96, 64), 1, ops, **kw),
                              f"nv12 {kw}")


def _yuv_frame(seed, fmt, w, h, pitch):
    """Random frame of the given format: every byte random, so the low 6 bits of the 10-bit words are noise too."""
    rng = np.random.default_rng(seed)
    rows = {_abi.CVGS_NV21: h + (h + 1) // 2, _abi.CVGS_P010: h + (h + 1) // 2, _abi.CVGS_P210: 2 * h, _abi.CVGS_Y210: h}[fmt]
    return rng.integers(0, 256, size=(rows, pitch), dtype=np.uint8)


FORMATS = [_abi.CVGS_NV21, _abi.CVGS_P010, _abi.CVGS_P210, _abi.CVGS_Y210]


@pytest.mark.parametrize("standard", [0, 3])
@pytest.mark.parametrize("fmt", FORMATS)
def test_other_yuv_formats_match_reference_kernel_and_oracle(fmt, standard):
    """ReadYUV<NV21 / P010 / P210 / Y210> + ConvertYUVToRGB (reference color_conversion.cuh:296-345,235-291)."""
    if gpu_util.fkref_lib(16) is None:
        pytest.skip("oracle/_ref not built")
    pitch = 2048
    sizes = [(320, 240), (322, 242), (162, 122), (64, 36), (2, 2)]
    frames = [_yuv_frame(1000 + 10 * standard + i + fmt, fmt, w, h, pitch) for i, (w, h) in enumerate(sizes)]
    for dsize in [(64, 128), (400, 300), (33, 7)]:
        ours = _ours(frames, sizes, pitch, dsize, standard, OPS, src_type=fmt)
        orc = _oracle(frames, sizes, pitch, dsize, standard, OPS, src_type=fmt)
        util.assert_bit_equal(ours, orc, f"format {fmt:#x} standard {standard} dsize {dsize}: ours vs oracle")
        for i, (f, (w, h)) in enumerate(zip(frames, sizes)):
            ref = gpu_util.run_fkref_yuv(fmt, f, w, h, dsize, standard, MUL, SUB, DIV)
            util.assert_bit_equal(ours[i], ref, f"format {fmt:#x} standard {standard} dsize {dsize} frame {i}: ours vs reference kernel")


@pytest.mark.parametrize("fmt", FORMATS)
def test_other_yuv_formats_modes(fmt):
    pitch = 1280
    sizes = [(300, 200), (128, 128)]
    frames = [_yuv_frame(1100 + i + fmt, fmt, w, h, pitch) for i, (w, h) in enumerate(sizes)]
    ops = [("reorder", (2, 1, 0)), ("mul", (0.5, 0.25, 2.0)), ("add", (1.0, 2.0, 3.0))]
    for standard, kw in [(1, dict(aspect=_abi.PRESERVE_AR, background=(1.0, 2.0, 3.0))), (2, dict(fp_contract=_abi.FP_SEPARATE)),
                         (0, dict(layout=_abi.OUT_NHWC))]:
        got = _ours(frames, sizes, pitch, (96, 64), standard, ops, src_type=fmt, **kw)
        want = _oracle(frames, sizes, pitch, (96, 64), standard, ops, src_type=fmt, **kw)
        util.assert_bit_equal(got.reshape(-1), want.reshape(-1), f"format {fmt:#x} {kw}")


def test_yuv_alignment_is_checked():
    lib = _abi.load()
    t = torch.zeros(1 << 16, dtype=torch.uint8, device="cuda")
    out = torch.zeros(3 * 8 * 8, device="cuda")
    crops = (_abi.Crop * 1)()
    crops[0].data, crops[0].width, crops[0].height, crops[0].pitch = t.data_ptr() + 2, 16, 16, 128
    p = util.make_pipeline((8, 8), [], out_ptr=out.data_ptr(), src_type=_abi.CVGS_Y210)
    assert lib.cvgs_b200_preproc_launch(crops, 1, 1, C.byref(p), None) != 0 and b"aligned" in lib.cvgs_b200_last_error()


@pytest.mark.parametrize("seed", range(20 + int(__import__("os").environ.get("CVGS_FUZZ_EXTRA", "0"))))
def test_random_yuv_frames_against_oracle(seed):
    """Every format, odd and even frame sizes (the last column / row shares its chroma sample), random pitches, sizes,
    standards, aspect modes and chains."""
    rng = np.random.default_rng(12000 + seed)
    fmt = [_abi.CVGS_NV12] + FORMATS
    fmt = fmt[seed % len(fmt)]
    n = int(rng.integers(1, 6))
    sizes = [(int(rng.integers(1, 140)), int(rng.integers(1, 100))) for _ in range(n)]
    wmax = max(w for w, _ in sizes)
    per_px = {_abi.CVGS_NV12: 1, _abi.CVGS_NV21: 1, _abi.CVGS_P010: 2, _abi.CVGS_P210: 2, _abi.CVGS_Y210: 4}[fmt]
    pitch = ((wmax + 1) // 2 * 2 * per_px + 7) // 8 * 8 + 8 * int(rng.integers(0, 4))
    frames = []
    for (w, h) in sizes:
        rows = {_abi.CVGS_NV12: h + (h + 1) // 2, _abi.CVGS_NV21: h + (h + 1) // 2, _abi.CVGS_P010: h + (h + 1) // 2,
                _abi.CVGS_P210: 2 * h, _abi.CVGS_Y210: h}[fmt]
        frames.append(rng.integers(0, 256, size=(rows, pitch), dtype=np.uint8))
    dsize = (int(rng.integers(1, 120)), int(rng.integers(1, 90)))
    ops = [OPS, [], [("reorder", (2, 1, 0)), ("mul", (0.5, 0.25, 2.0))], [("gray", (1,)), ("sub", (3.0,))]][int(rng.integers(0, 4))]
    kw = dict(aspect=int(rng.integers(0, 4)), background=(1.0, 2.0, 3.0))
    standard = int(rng.integers(0, 4))
    lib = _abi.load()
    d = [torch.from_numpy(f).cuda() for f in frames]
    crops_d, crops_h = (_abi.Crop * n)(), (_abi.Crop * n)()
    for i, (t, f, (w, h)) in enumerate(zip(d, frames, sizes)):
        crops_d[i].data, crops_d[i].width, crops_d[i].height, crops_d[i].pitch = t.data_ptr(), w, h, pitch
        crops_h[i].data, crops_h[i].width, crops_h[i].height, crops_h[i].pitch = f.ctypes.data, w, h, pitch
    nco = util.out_channels(_abi.CVGS_8UC3, ops)
    out = torch.full((n, nco, dsize[1], dsize[0]), float("nan"), device="cuda")
    want = np.full((n, nco, dsize[1], dsize[0]), np.nan, dtype=np.float32)
    p = util.make_pipeline(dsize, ops, out_ptr=out.data_ptr(), src_type=fmt, yuv_standard=standard, **kw)
    _abi.check(lib.cvgs_b200_preproc_launch(crops_d, n, n, C.byref(p), None))
    torch.cuda.synchronize()
    p = util.make_pipeline(dsize, ops, out_ptr=want.ctypes.data, src_type=fmt, yuv_standard=standard, **kw)
    assert util.oracle_lib().oracle_preproc(crops_h, n, n, C.byref(p), 0) == 0
    util.assert_bit_equal(out.cpu().numpy(), want, f"seed {seed} fmt {fmt:#x} sizes {sizes} pitch {pitch} dsize {dsize} std {standard} ops {ops} {kw}")


@pytest.mark.parametrize("src_type", [_abi.CVGS_NV12, _abi.CVGS_NV21])
@pytest.mark.parametrize("standard", [0, 1, 2, 3])
def test_tma_staged_kernel_takes_nv12_and_nv21(src_type, standard):
    """Batches of even-sized NV12 / NV21 frames in the common geometry go through the TMA-staged kernel
    (preproc_yuv_tma.cuh; forced: variant 2 fails instead of falling back): luma and chroma planes staged through their own
    tensor maps, conversion of every tap in the scaled domain.  Up- and down-scales, frame bases at every 16-byte phase,
    several destination sizes (one / several column bands, odd heights), the FMA-DIV chain, a generic chain and no chain;
    against the oracle, and against the reference's own kernel for NV12."""
    lib = _abi.load()
    pitch = 512
    sizes = [(320, 240), (322, 242), (160, 120), (64, 36), (2, 2), (500, 300), (96, 400)]
    frames = [_frame(70 * standard + i, w, h, pitch) for i, (w, h) in enumerate(sizes)]
    prev = lib.cvgs_b200_set_kernel_variant(2)
    try:
        for dsize, ops in [((64, 128), OPS), ((400, 300), OPS), ((33, 7), [("mul", (0.5, 0.25, 2.0)), ("add", (1.0, 2.0, 3.0))]),
                           ((224, 225), [])]:
            ours = _ours(frames, sizes, pitch, dsize, standard, ops, src_type=src_type)
            orc = _oracle(frames, sizes, pitch, dsize, standard, ops, src_type=src_type)
            util.assert_bit_equal(ours, orc, f"standard {standard} dsize {dsize}: TMA-staged kernel vs oracle")
            if src_type == _abi.CVGS_NV12 and ops is OPS and gpu_util.fkref_lib(16) is not None:
                for i, (f, (w, h)) in enumerate(zip(frames, sizes)):
                    ref = gpu_util.run_fkref_nv12(f, w, h, dsize, standard, MUL, SUB, DIV)
                    util.assert_bit_equal(ours[i], ref, f"standard {standard} dsize {dsize} frame {i}: vs reference kernel")
        # frames at odd byte offsets of their buffer (plane bases at other 16-byte phases)
        n = 5
        big = torch.from_numpy(np.random.default_rng(5).integers(0, 256, size=(n * 400 * pitch + 64,), dtype=np.uint8)).cuda()
        host = big.cpu().numpy()
        crops, ocrops = (_abi.Crop * n)(), (_abi.Crop * n)()
        for i in range(n):
            off = i * 400 * pitch + 3 * i + 1
            crops[i].data, crops[i].width, crops[i].height, crops[i].pitch = big.data_ptr() + off, 200, 120, pitch
            ocrops[i].data, ocrops[i].width, ocrops[i].height, ocrops[i].pitch = host.ctypes.data + off, 200, 120, pitch
        out = torch.full((n, 3, 60, 100), float("nan"), device="cuda")
        want = np.full((n, 3, 60, 100), np.nan, dtype=np.float32)
        p = util.make_pipeline((100, 60), OPS, out_ptr=out.data_ptr(), src_type=src_type, yuv_standard=standard)
        po = util.make_pipeline((100, 60), OPS, out_ptr=want.ctypes.data, src_type=src_type, yuv_standard=standard)
        _abi.check(lib.cvgs_b200_preproc_launch(crops, n, n, C.byref(p), None))
        torch.cuda.synchronize()
        assert util.oracle_lib().oracle_preproc(ocrops, n, n, C.byref(po), 0) == 0
        util.assert_bit_equal(out.cpu().numpy(), want, "unaligned plane bases")
    finally:
        lib.cvgs_b200_set_kernel_variant(prev)


@pytest.mark.parametrize("src_type", [_abi.CVGS_P010, _abi.CVGS_P210])
@pytest.mark.parametrize("standard", [0, 1, 2, 3])
def test_tma_staged_kernel_takes_p010_and_p210(src_type, standard):
    """The 16-bit semi-planar formats through the TMA-staged kernel (DEPTH = 2 instantiation; forced: variant 2 fails
    instead of falling back): halfword samples with noise in their low six bits, chroma pairs as aligned words, 4:2:0 and
    4:2:2 chroma rows.  Against the oracle, and for the matrices the reference harness holds (0 and 3) against the
    reference's own kernel."""
    lib = _abi.load()
    pitch = 1024
    sizes = [(320, 240), (322, 242), (160, 120), (64, 36), (2, 2), (500, 300), (96, 400)]
    frames = [_yuv_frame(170 * standard + i + src_type, src_type, w, h, pitch) for i, (w, h) in enumerate(sizes)]
    prev = lib.cvgs_b200_set_kernel_variant(2)
    try:
        for dsize, ops in [((64, 128), OPS), ((400, 300), OPS), ((33, 7), [("mul", (0.5, 0.25, 2.0)), ("add", (1.0, 2.0, 3.0))]),
                           ((224, 225), [])]:
            ours = _ours(frames, sizes, pitch, dsize, standard, ops, src_type=src_type)
            orc = _oracle(frames, sizes, pitch, dsize, standard, ops, src_type=src_type)
            util.assert_bit_equal(ours, orc, f"format {src_type:#x} standard {standard} dsize {dsize}: TMA-staged kernel vs oracle")
            if standard in (0, 3) and ops is OPS and gpu_util.fkref_lib(16) is not None:
                for i, (f, (w, h)) in enumerate(zip(frames, sizes)):
                    ref = gpu_util.run_fkref_yuv(src_type, f, w, h, dsize, standard, MUL, SUB, DIV)
                    util.assert_bit_equal(ours[i], ref, f"format {src_type:#x} standard {standard} dsize {dsize} frame {i}: vs reference kernel")
        # plane bases at the other 4-byte phases of a 16-byte line
        n = 4
        rows = 400
        big = torch.from_numpy(np.random.default_rng(6).integers(0, 256, size=(n * rows * pitch + 64,), dtype=np.uint8)).cuda()
        host = big.cpu().numpy()
        crops, ocrops = (_abi.Crop * n)(), (_abi.Crop * n)()
        for i in range(n):
            off = i * rows * pitch + 4 * i
            crops[i].data, crops[i].width, crops[i].height, crops[i].pitch = big.data_ptr() + off, 200, 120, pitch
            ocrops[i].data, ocrops[i].width, ocrops[i].height, ocrops[i].pitch = host.ctypes.data + off, 200, 120, pitch
        out = torch.full((n, 3, 60, 100), float("nan"), device="cuda")
        want = np.full((n, 3, 60, 100), np.nan, dtype=np.float32)
        p = util.make_pipeline((100, 60), OPS, out_ptr=out.data_ptr(), src_type=src_type, yuv_standard=standard)
        po = util.make_pipeline((100, 60), OPS, out_ptr=want.ctypes.data, src_type=src_type, yuv_standard=standard)
        _abi.check(lib.cvgs_b200_preproc_launch(crops, n, n, C.byref(p), None))
        torch.cuda.synchronize()
        assert util.oracle_lib().oracle_preproc(ocrops, n, n, C.byref(po), 0) == 0
        util.assert_bit_equal(out.cpu().numpy(), want, "plane bases at other word phases")
        # a frame at a 2-byte phase is not this kernel's (chroma pairs must be aligned words): declined, not wrong
        crops[0].data = big.data_ptr() + 2
        assert lib.cvgs_b200_preproc_launch(crops, 1, 1, C.byref(p), None) != 0
    finally:
        lib.cvgs_b200_set_kernel_variant(prev)


@pytest.mark.parametrize("standard", [0, 1, 2, 3])
def test_tma_staged_kernel_takes_y210(standard):
    """Packed Y210 frames through the TMA-staged kernel (PACKED instantiation; forced: variant 2): a tap is the 8-byte
    group of its pixel pair, Y at halfword 0 or 2, U and V at halfwords 1 and 3, noise in the low six bits of every word.
    Against the oracle, and for the matrices the reference harness holds against the reference's own kernel."""
    lib = _abi.load()
    src_type = _abi.CVGS_Y210
    pitch = 2048
    sizes = [(320, 240), (322, 242), (160, 120), (64, 36), (2, 2), (500, 300), (96, 400)]
    frames = [_yuv_frame(190 * standard + i, src_type, w, h, pitch) for i, (w, h) in enumerate(sizes)]
    prev = lib.cvgs_b200_set_kernel_variant(2)
    try:
        for dsize, ops in [((64, 128), OPS), ((400, 300), OPS), ((33, 7), [("mul", (0.5, 0.25, 2.0)), ("add", (1.0, 2.0, 3.0))]),
                           ((224, 225), [])]:
            ours = _ours(frames, sizes, pitch, dsize, standard, ops, src_type=src_type)
            orc = _oracle(frames, sizes, pitch, dsize, standard, ops, src_type=src_type)
            util.assert_bit_equal(ours, orc, f"Y210 standard {standard} dsize {dsize}: TMA-staged kernel vs oracle")
            if standard in (0, 3) and ops is OPS and gpu_util.fkref_lib(16) is not None:
                for i, (f, (w, h)) in enumerate(zip(frames, sizes)):
                    ref = gpu_util.run_fkref_yuv(src_type, f, w, h, dsize, standard, MUL, SUB, DIV)
                    util.assert_bit_equal(ours[i], ref, f"Y210 standard {standard} dsize {dsize} frame {i}: vs reference kernel")
        # frames at the other 8-byte phase of a 16-byte line
        n = 3
        rows = 200
        big = torch.from_numpy(np.random.default_rng(7).integers(0, 256, size=(n * rows * pitch + 64,), dtype=np.uint8)).cuda()
        host = big.cpu().numpy()
        crops, ocrops = (_abi.Crop * n)(), (_abi.Crop * n)()
        for i in range(n):
            off = i * rows * pitch + 8 * i
            crops[i].data, crops[i].width, crops[i].height, crops[i].pitch = big.data_ptr() + off, 200, 120, pitch
            ocrops[i].data, ocrops[i].width, ocrops[i].height, ocrops[i].pitch = host.ctypes.data + off, 200, 120, pitch
        out = torch.full((n, 3, 60, 100), float("nan"), device="cuda")
        want = np.full((n, 3, 60, 100), np.nan, dtype=np.float32)
        p = util.make_pipeline((100, 60), OPS, out_ptr=out.data_ptr(), src_type=src_type, yuv_standard=standard)
        po = util.make_pipeline((100, 60), OPS, out_ptr=want.ctypes.data, src_type=src_type, yuv_standard=standard)
        _abi.check(lib.cvgs_b200_preproc_launch(crops, n, n, C.byref(p), None))
        torch.cuda.synchronize()
        assert util.oracle_lib().oracle_preproc(ocrops, n, n, C.byref(po), 0) == 0
        util.assert_bit_equal(out.cpu().numpy(), want, "frames at the other 8-byte phase")
    finally:
        lib.cvgs_b200_set_kernel_variant(prev)
