"""The widened forms of the chain on one geometry (256 crops / frames -> 224x224, BASELINE config 3 shape): time per
launch, algorithmic bytes (tapped source bytes + output bytes), achieved GB/s and fraction of the measured HBM peak.
    python scripts/run_forms.py [reps]
Every form is spot-checked against the oracle before it is timed."""
import ctypes as C, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from cvgpuspeedup_b200 import _abi
from tests import util

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
lib = _abi.load()
try:
    PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    PEAK = 6454.6
N, DST = 256, (224, 224)
FW, FH = 3840, 2160
NORM = [("mul", (1 / 255.0,) * 3), ("sub", (0.485, 0.456, 0.406)), ("div", (0.229, 0.224, 0.225))]
NORM4 = [("mul", (1 / 255.0,) * 4), ("sub", (0.485, 0.456, 0.406, 0.5)), ("div", (0.229, 0.224, 0.225, 0.25))]
rng = np.random.default_rng(3)
rects = []
for _ in range(N):
    w, h = int(rng.integers(224, 897)), int(rng.integers(224, 897))
    rects.append((int(rng.integers(0, FW - w + 1)), int(rng.integers(0, FH - h + 1)), w, h))


def tapped_bytes(px_bytes):
    """bytes_in of SURVEY 8(d): distinct source pixels that are a tap of some output pixel, times the pixel size."""
    mask = np.zeros((FH, FW), dtype=bool)
    for (x, y, w, h) in rects:
        fx = np.float32(1.0 / (np.float64(DST[0]) / w)); fy = np.float32(1.0 / (np.float64(DST[1]) / h))
        xs = np.floor(np.arange(DST[0], dtype=np.float32) * fx).astype(int); ys = np.floor(np.arange(DST[1], dtype=np.float32) * fy).astype(int)
        xi = np.unique(np.concatenate([xs, np.minimum(xs + 1, w - 1)])) + x
        yi = np.unique(np.concatenate([ys, np.minimum(ys + 1, h - 1)])) + y
        mask[np.ix_(yi, xi)] = True
    return int(mask.sum()) * px_bytes


def run(name, src_type, ops, px_bytes, image, crops_of, out_channels=3, out_bytes_per_value=4, variant=0, src_bytes=None, **kw):
    d_img = torch.from_numpy(image).cuda()
    n_out = N * out_channels * DST[0] * DST[1]
    if out_bytes_per_value == 1:
        out = torch.empty(n_out, dtype=torch.uint8, device="cuda")
    else:
        out = torch.empty(n_out, dtype=torch.float32, device="cuda")
    crops = crops_of(d_img.data_ptr())
    p = util.make_pipeline(DST, ops, out_ptr=out.data_ptr(), src_type=src_type, **kw)
    prev = lib.cvgs_b200_set_kernel_variant(variant)
    st = torch.cuda.current_stream().cuda_stream
    par = (_abi.Parent * N)()  # the frame every crop was cut from (GpuMat::datastart / locateROI)
    for i in range(N):
        par[i].datastart, par[i].whole_width, par[i].whole_height = d_img.data_ptr(), FW, FH
    use_par = px_bytes in (3, 4, 6, 8)

    def launch():
        if use_par:
            return lib.cvgs_b200_preproc_launch_ex(crops, par, N, N, C.byref(p), st)
        return lib.cvgs_b200_preproc_launch(crops, N, N, C.byref(p), st)
    for _ in range(3):
        _abi.check(launch())
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        launch()
    e1.record()
    torch.cuda.synchronize()
    lib.cvgs_b200_set_kernel_variant(prev)
    us = e0.elapsed_time(e1) * 1e3 / reps
    # spot check: three planes against the oracle
    idx = [0, 100, 255]
    hc = crops_of(image.ctypes.data)
    sub = (_abi.Crop * 3)(*[hc[i] for i in idx])
    plane = out_channels * DST[0] * DST[1]
    want = np.zeros(3 * plane, dtype=np.uint8 if out_bytes_per_value == 1 else np.float32)
    po = util.make_pipeline(DST, ops, out_ptr=want.ctypes.data, src_type=src_type, **kw)
    assert util.oracle_lib().oracle_preproc(sub, 3, 3, C.byref(po), 0) == 0
    got = out.cpu().numpy().reshape(N, plane)[idx].reshape(-1)
    assert np.array_equal(got.view(np.uint8), want.view(np.uint8)), f"{name}: differs from the oracle"
    b_in = tapped_bytes(px_bytes) if px_bytes else int(src_bytes or 0)
    b_out = n_out * out_bytes_per_value
    gbs = (b_in + b_out) / us / 1e3
    print(f"{name:<46} {us:8.1f} us  in {b_in / 1e6:6.1f} MB  out {b_out / 1e6:6.1f} MB  {gbs:7.0f} GB/s  {gbs / PEAK:5.2f} of HBM peak", flush=True)


def rgb_crops(px):
    def f(base):
        c = (_abi.Crop * N)()
        for i, (x, y, w, h) in enumerate(rects):
            c[i].data, c[i].width, c[i].height, c[i].pitch = base + y * pitch + px * x, w, h, pitch
        return c
    return f


print(f"256 crops (224..896 px) of a {FW}x{FH} frame -> 224x224, one launch; peak {PEAK} GB/s")
pitch = 3 * FW
img3 = rng.integers(0, 256, size=(FH, pitch), dtype=np.uint8)
run("CV_8UC3 -> NCHW float, TMA-staged kernel", _abi.CVGS_8UC3, [("reorder", (2, 1, 0))] + NORM, 3, img3, rgb_crops(3))
run("CV_8UC3 -> NCHW float, direct-gather kernel", _abi.CVGS_8UC3, [("reorder", (2, 1, 0))] + NORM, 3, img3, rgb_crops(3), variant=1)
run("CV_8UC3 -> CV_8UC3 (convertTo 8U, packed)", _abi.CVGS_8UC3, [], 3, img3, rgb_crops(3), out_bytes_per_value=1,
    layout=_abi.OUT_NHWC, dst_type=_abi.CVGS_8UC3)
run("CV_8UC3 -> gray (BGR2GRAY), 1 plane", _abi.CVGS_8UC3, [("reorder", (2, 1, 0)), ("gray", (0,)), ("mul", (1 / 255.0,))], 3, img3,
    rgb_crops(3), out_channels=1)
run("CV_8UC3 -> RGBA float (BGR2RGBA)", _abi.CVGS_8UC3, [("reorder", (2, 1, 0)), ("add_alpha", (255.0,))] + NORM4, 3, img3, rgb_crops(3),
    out_channels=4)
pitch = 6 * FW
img6 = rng.integers(0, 256, size=(FH, pitch), dtype=np.uint8)
run("CV_16UC3 -> NCHW float", _abi.CVGS_16UC3, NORM, 6, img6, rgb_crops(6))
run("CV_16SC3 -> NCHW float", _abi.CVGS_16SC3, NORM, 6, img6, rgb_crops(6))
pitch = 4 * FW
img4 = rng.integers(0, 256, size=(FH, pitch), dtype=np.uint8)
run("CV_8UC4 -> NCHW float (4 planes)", _abi.CVGS_8UC4, NORM4, 4, img4, rgb_crops(4), out_channels=4)

# YUV frames: a crop is a whole frame, so 256 frames of 448x448 (2x down-scale) stand in for the crops
fw = fh = 448
rects = [(0, 0, fw, fh)] * N


def yuv_frames(rows, bytes_per_luma):
    pitch_y = 1024 if bytes_per_luma <= 2 else 2048
    frames = rng.integers(0, 256, size=(8, rows, pitch_y), dtype=np.uint8)

    def f(base):
        c = (_abi.Crop * N)()
        for i in range(N):
            c[i].data, c[i].width, c[i].height, c[i].pitch = base + (i % 8) * rows * pitch_y, fw, fh, pitch_y
        return c
    return frames.reshape(8 * rows, pitch_y), f


FW, FH = fw, fh
for name, fmt, rows, bpl, bpp in [("NV12 frames 448x448 -> NCHW float", _abi.CVGS_NV12, fh + fh // 2, 1, 1.5),
                                  ("P010 frames 448x448 -> NCHW float", _abi.CVGS_P010, fh + fh // 2, 2, 3.0),
                                  ("Y210 frames 448x448 -> NCHW float", _abi.CVGS_Y210, fh, 4, 4.0)]:
    image, cf = yuv_frames(rows, bpl)
    d_img = None
    # source bytes: every sample of the 8 distinct frames is tapped at a 2x down-scale (both rows and columns of a 2x2 cell)
    run(name, fmt, NORM, 0, image, cf, yuv_standard=1, src_bytes=8 * fw * fh * bpp)
