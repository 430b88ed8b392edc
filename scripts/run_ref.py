"""Runs the reference's own kernels (oracle/_ref, built from /root/reference/fkl/include) on the bench workloads, for ncu
captures next to ours:   python scripts/run_ref.py c2|c3|c4 [reps]"""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import bench

case = sys.argv[1] if len(sys.argv) > 1 else "c3"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
f3 = lambda v: (C.c_float * 3)(*v)  # noqa: E731
s = torch.cuda.current_stream()


def batch_lib(n):
    lib = C.CDLL(os.path.join(ROOT, "oracle", "_ref", f"libfkref_{n}.so"))
    fn = getattr(lib, f"fkref_preproc_{n}")
    fn.restype = C.c_int
    fn.argtypes = [C.POINTER(C.c_void_p), C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int), C.c_int, C.c_int, C.c_int,
                   C.c_int, C.POINTER(C.c_float), C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_float), C.POINTER(C.c_float),
                   C.c_void_p, C.c_void_p]
    return fn


if case == "c2":
    fn = batch_lib(50)
    frames = bench.make_frames(4, seed=2)
    for img, rects in frames * reps:
        d = torch.from_numpy(img).cuda()
        out = torch.empty((50, 3, 128, 64), device="cuda")
        ptrs = (C.c_void_p * 50)(*[d.data_ptr() + y * bench.PITCH + 3 * x for (x, y, w, h) in rects])
        rc = fn(ptrs, (C.c_int * 50)(*[r[2] for r in rects]), (C.c_int * 50)(*[r[3] for r in rects]), (C.c_int * 50)(*[bench.PITCH] * 50),
                50, 64, 128, 1, f3((0, 0, 0)), 1, f3(bench.MUL), f3(bench.SUB), f3(bench.DIV), out.data_ptr(), s.cuda_stream)
        assert rc == 0
        torch.cuda.synchronize()
elif case == "c3":
    fn = batch_lib(128)
    img, rects = bench.make_c3(seed=3)
    d = torch.from_numpy(img).cuda()
    out = torch.empty((256, 3, 224, 224), device="cuda")
    for _ in range(reps):
        for half in range(2):
            rr = rects[128 * half:128 * half + 128]
            ptrs = (C.c_void_p * 128)(*[d.data_ptr() + y * img.shape[1] + 3 * x for (x, y, w, h) in rr])
            rc = fn(ptrs, (C.c_int * 128)(*[r[2] for r in rr]), (C.c_int * 128)(*[r[3] for r in rr]), (C.c_int * 128)(*[img.shape[1]] * 128),
                    128, 224, 224, 1, f3((0, 0, 0)), 1, f3((1 / 255.0,) * 3), f3((0.485, 0.456, 0.406)), f3((0.229, 0.224, 0.225)),
                    out[128 * half:].data_ptr(), s.cuda_stream)
            assert rc == 0
        torch.cuda.synchronize()
else:
    ref = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libfkref_ct.so"))
    ref.fkref_ct_create.restype = C.c_void_p
    ref.fkref_ct_create.argtypes = [C.c_int] * 4
    ref.fkref_ct_update.restype = C.c_int
    ref.fkref_ct_update.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_float),
                                    C.POINTER(C.c_float), C.c_void_p]
    h = ref.fkref_ct_create(16, 0, 1920, 1080)
    rng = np.random.default_rng(4)
    frame = torch.from_numpy(rng.integers(0, 256, size=(1080, 6144), dtype=np.uint8)).cuda()
    for _ in range(16 + reps):
        assert ref.fkref_ct_update(h, frame.data_ptr(), 1920, 1080, 6144, 1, f3((1 / 255.0,) * 3), f3((0.485, 0.456, 0.406)),
                                   f3((0.229, 0.224, 0.225)), s.cuda_stream) == 0
        torch.cuda.synchronize()
print("done", case)
