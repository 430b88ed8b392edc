"""Host-side cost breakdown of the small-batch launch path (cvgs_b200_debug_host_profile) on the c2 workload."""
import ctypes as C, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from cvgpuspeedup_b200 import _abi
from tests import util
lib = _abi.load()
nsets = 32
sets = []
for k in range(nsets):
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
n_arr = (C.c_int32 * nsets)(*[n] * nsets)
def seq(k, parents):
    if parents:
        _abi.check(lib.cvgs_b200_preproc_launch_sequence_ex(crops_pp, par_pp, n_arr, n_arr, pipes_pp, nsets, k, st.cuda_stream))
    else:
        _abi.check(lib.cvgs_b200_preproc_launch_sequence(crops_pp, n_arr, n_arr, pipes_pp, nsets, k, st.cuda_stream))
st = torch.cuda.Stream()
lib.cvgs_b200_debug_host_profile.argtypes = [C.POINTER(C.c_double), C.c_int]
for variant, steps, parents in ((0, 3200, 0), (0, 400, 0), (0, 3200, 1), (0, 400, 1), (1, 400, 0)):
    lib.cvgs_b200_set_kernel_variant(variant)
    seq(64, parents)
    torch.cuda.synchronize()
    out5 = (C.c_double * 5)()
    lib.cvgs_b200_debug_host_profile(out5, 1)
    t0 = time.perf_counter()
    seq(steps, parents)
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    lib.cvgs_b200_debug_host_profile(out5, 1)
    c = max(out5[0], 1)
    print(f"variant {variant} parents {parents}: host issue {1e6*(t1-t0)/steps:.2f} us/launch over {steps} launches, total {1e6*(t2-t0)/steps:.2f} us/launch; "
          f"profile calls {out5[0]:.0f}: fill {out5[1]/c:.2f} plan {out5[2]/c:.2f} encode {out5[3]/c:.2f} launch {out5[4]/c:.2f} us")
