"""Frame-loop probe: host time of cvgs_b200_preproc_launch_sequence_ex (wall clock until the call returns) against the
device time of the same steps (CUDA events), on the bench workload.  Tells a host-bound loop from a GPU-bound one.
  python scripts/seq_probe.py [steps]"""
import ctypes as C, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from cvgpuspeedup_b200 import _abi, api

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 100
F = 32
lib = _abi.load()
frames = bench.make_frames(F, seed=2)
sets = bench.Frames(torch, frames)
stream = torch.cuda.Stream()
sp = stream.cuda_stream
lib.cvgs_b200_set_overlap(1)
alg = sum(sum(bench.algorithmic_bytes(r)) for _, r in frames) / F
for label, coalesce in (("shared launches", 1), ("one launch per frame, helper threads", 0)):
    lib.cvgs_b200_set_coalesce(coalesce)
    sets.device_steps(lib, 200, sp)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0 = lib.cvgs_b200_launch_count()
    e0.record(stream)
    t0 = time.perf_counter()
    sets.device_steps(lib, steps, sp)
    t1 = time.perf_counter()
    e1.record(stream)
    torch.cuda.synchronize()
    n = steps * F
    dev_us = e0.elapsed_time(e1) * 1e3 / n
    print(f"{label}: {lib.cvgs_b200_launch_count() - l0} launches for {n} frames; host {1e6 * (t1 - t0) / n:.3f} us/frame, "
          f"device {dev_us:.3f} us/frame = {alg / dev_us / 1e3:.0f} GB/s", flush=True)
