"""Run one workload repeatedly through the C-ABI (for ncu captures and quick timing):
   python scripts/run_case.py c2|c3 [reps] [variant] [overlap] [parents]"""
import ctypes as C, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from cvgpuspeedup_b200 import _abi
from tests import util
case = sys.argv[1] if len(sys.argv) > 1 else "c3"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
variant = int(sys.argv[3]) if len(sys.argv) > 3 else 0
lib = _abi.load()
lib.cvgs_b200_set_kernel_variant(variant)
overlap = int(sys.argv[4]) if len(sys.argv) > 4 else 0
lib.cvgs_b200_set_overlap(overlap)
nsets = 2 if case == "c3" else 32
sets = []
for k in range(nsets):
    # C3_LO / C3_HI: crop-size range of the c3 workload (default 224..896); C3_FORCE_BIG=1 makes crop 0 an 896x896 one, so
    # that workloads of one scale run with the ring geometry (and residency) of the mixed workload
    if case == "c3":
        w = util.workload_c3(seed=3 + k, lo=int(os.environ.get("C3_LO", "224")), hi=int(os.environ.get("C3_HI", "896")))
        if os.environ.get("C3_FORCE_BIG"):
            w.rects[0] = (100, 100, 896, 896)
    else:
        w = util.workload_c2(seed=2 + k)
    d_img = torch.from_numpy(w.image).cuda()
    d_out = torch.empty((len(w.rects), 3, w.dsize[1], w.dsize[0]), dtype=torch.float32, device="cuda")
    sets.append((w, d_img, d_out, util.host_crops(w.image, w.rects, base_ptr=d_img.data_ptr()),
                 util.make_pipeline(w.dsize, w.ops, out_ptr=d_out.data_ptr()),
                 util.host_parents(w.image, w.width, w.height, len(w.rects), base_ptr=d_img.data_ptr())))
n = len(sets[0][0].rects)
crops_pp = (C.POINTER(_abi.Crop) * nsets)(*[C.cast(s[3], C.POINTER(_abi.Crop)) for s in sets])
pipes_pp = (C.POINTER(_abi.Pipeline) * nsets)(*[C.pointer(s[4]) for s in sets])
par_pp = (C.POINTER(_abi.Parent) * nsets)(*[C.cast(s[5], C.POINTER(_abi.Parent)) for s in sets])
use_parents = int(sys.argv[5]) if len(sys.argv) > 5 else 1
n_arr = (C.c_int32 * nsets)(*[n] * nsets)
st = torch.cuda.Stream()
def run(k):
    if use_parents:
        _abi.check(lib.cvgs_b200_preproc_launch_sequence_ex(crops_pp, par_pp, n_arr, n_arr, pipes_pp, nsets, k, st.cuda_stream))
    else:
        _abi.check(lib.cvgs_b200_preproc_launch_sequence(crops_pp, n_arr, n_arr, pipes_pp, nsets, k, st.cuda_stream))
run(nsets); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.perf_counter(); e0.record(st); run(reps * nsets); t1 = time.perf_counter(); e1.record(st); torch.cuda.synchronize()
us = e0.elapsed_time(e1) * 1e3 / (reps * nsets)
print(f"{case} variant {variant} overlap {overlap} parents {use_parents}: {us:.2f} us/launch device, host issue {1e6*(t1-t0)/(reps*nsets):.2f} us/launch, {n/us:.3f} Mcrops/s")
w0 = sets[0][0]
idx = list(range(0, n, max(1, n // 6)))
want = util.run_oracle(w0.image, [w0.rects[i] for i in idx], w0.dsize, w0.ops)
util.assert_bit_equal(sets[0][2][idx].cpu().numpy(), want, "spot check vs oracle")
print("parity ok")
