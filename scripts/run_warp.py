"""Batched perspective/affine warp timing: python scripts/run_warp.py [batch] [dst_w dst_h] [reps]
Sources are 1920x1080 CV_8UC3 images; every plane gets its own perspective matrix (a jittered quadrilateral mapped onto
the destination), then Mul/Sub/Div and the planar split.  The reference's kernel is timed next to it, one image per
launch (its batch size is a template parameter; oracle/_ref/libfkref_16.so holds the single-image instantiation)."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from cvgpuspeedup_b200 import _abi
import cvgpuspeedup_b200 as cvGS
from tests import util
from tests.test_warp_gpu import perspective_from_points
n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
W = int(sys.argv[2]) if len(sys.argv) > 2 else 224
H = int(sys.argv[3]) if len(sys.argv) > 3 else 224
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 50
lib = _abi.load()
rng = np.random.default_rng(4)
SW, SH, PITCH = 1920, 1080, 6144
frames = [torch.from_numpy(util.make_image(rng, SW, SH, pitch=PITCH)).cuda() for _ in range(4)]
crops = (_abi.Crop * n)()
warps = (_abi.Warp * n)()
invs = []
for i in range(n):
    f = frames[i % 4]
    crops[i].data, crops[i].width, crops[i].height, crops[i].pitch = f.data_ptr(), SW, SH, PITCH
    cx, cy, s = rng.uniform(400, 1500), rng.uniform(300, 800), rng.uniform(120, 300)
    j = lambda: rng.uniform(-0.15, 0.15) * s
    src = [(cx - s + j(), cy - s + j()), (cx + s + j(), cy - s + j()), (cx - s + j(), cy + s + j()), (cx + s + j(), cy + s + j())]
    m = perspective_from_points(src, [(0, 0), (W, 0), (0, H), (W, H)])
    inv = cvGS.api.invert_warp_matrix(m, cvGS.WARP_PERSPECTIVE)
    invs.append(inv)
    warps[i].type = cvGS.WARP_PERSPECTIVE
    for k in range(9):
        warps[i].m[k] = float(inv[k])
out = torch.empty((n, 3, H, W), device="cuda")
ops = [("mul", (1 / 255.0,) * 3), ("sub", (0.485, 0.456, 0.406)), ("div", (0.229, 0.224, 0.225))]
p = util.make_pipeline((W, H), ops, out_ptr=out.data_ptr())
st = torch.cuda.current_stream().cuda_stream
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
out_bytes = 12 * W * H * n


def timed(pipe, label):
    for _ in range(5):
        _abi.check(lib.cvgs_b200_warp_launch(crops, warps, n, n, C.byref(pipe), st))
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        lib.cvgs_b200_warp_launch(crops, warps, n, n, C.byref(pipe), st)
    e1.record()
    torch.cuda.synchronize()
    t = e0.elapsed_time(e1) * 1e3 / reps
    print(f"warp batch {n} -> {W}x{H} [{label}]: {t:.1f} us/batch, {n / t * 1e6:.0f} planes/s, output stream {out_bytes / t / 1e3:.0f} GB/s", flush=True)
    return t


timed(p, "mul, sub, div, split")
us = timed(util.make_pipeline((W, H), ops[:1], out_ptr=out.data_ptr()), "mul, split")
prev = lib.cvgs_b200_set_kernel_variant(1)
timed(p, "general kernel: mul, sub, div, split")
lib.cvgs_b200_set_kernel_variant(prev)
path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", "libfkref_16.so")
if os.path.exists(path):
    ref = C.CDLL(path)
    fn = ref.fkref_warp_16
    fn.restype = C.c_int
    fn.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_float), C.c_int, C.c_int,
                   C.POINTER(C.c_float), C.c_void_p, C.c_int, C.c_void_p]
    ms = [(C.c_float * 9)(*[float(v) for v in inv]) for inv in invs]
    mul = (C.c_float * 3)(1 / 255.0, 1 / 255.0, 1 / 255.0)
    def ref_batch():
        for i in range(n):
            fn(1, 0, crops[i].data, SW, SH, PITCH, ms[i], W, H, mul, out[i].data_ptr(), 0, st)
    for _ in range(3):
        ref_batch()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(max(1, reps // 5)):
        ref_batch()
    e1.record()
    torch.cuda.synchronize()
    rus = e0.elapsed_time(e1) * 1e3 / max(1, reps // 5)
    print(f"reference kernel, {n} single-image launches (warp + Mul + split): {rus:.1f} us/batch ({us and rus / us:.1f}x)")
