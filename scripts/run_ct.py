"""CircularTensor update timing (BASELINE config 4): python scripts/run_ct.py [plane_w plane_h] [depth] [reps]
Frames are 1920x1080 CV_8UC3; the plane is the resize target (1920x1080 = no resize)."""
import ctypes as C, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from cvgpuspeedup_b200 import _abi
import cvgpuspeedup_b200 as cvGS
from tests import util
W = int(sys.argv[1]) if len(sys.argv) > 1 else 640
H = int(sys.argv[2]) if len(sys.argv) > 2 else 360
depth = int(sys.argv[3]) if len(sys.argv) > 3 else 16
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 50
lib = _abi.load()
rng = np.random.default_rng(4)
frames = [torch.from_numpy(util.make_image(rng, 1920, 1080, pitch=6144)).cuda() for _ in range(8)]
ct = cvGS.CircularTensor(W, H, depth, cvGS.CT_NEWEST_FIRST, cvGS.CT_STANDARD)
st = torch.cuda.Stream()
ops = [cvGS.cvtColor(cvGS.COLOR_BGR2RGB), cvGS.multiply((1 / 255.0,) * 3), cvGS.subtract((0.485, 0.456, 0.406)), cvGS.divide((0.229, 0.224, 0.225))]
mats = [cvGS.GpuMat(f.data_ptr(), 1920, 1080, 6144, owner=f) for f in frames]
for i in range(depth + 4):
    ct.update(st, mats[i % 8], *ops)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(st)
for i in range(reps):
    ct.update(st, mats[i % 8], *ops)
e1.record(st)
torch.cuda.synchronize()
us = e0.elapsed_time(e1) * 1e3 / reps
plane = 12 * W * H
src = 3 * 1920 * 1080 if (W, H) == (1920, 1080) else min(3 * 1920 * 1080, 4 * 3 * W * H)
alg = src + (depth - 1) * plane + depth * plane
print(f"CircularTensor {W}x{H} depth {depth}: {us:.1f} us/update, algorithmic {alg/1e6:.1f} MB -> {alg/us/1e3:.0f} GB/s")
