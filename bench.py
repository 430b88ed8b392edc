#!/usr/bin/env python
"""bench.py -- crops/s of the fused crop -> resize -> normalise -> split path on B200.

Contract (see the task statement): `python bench.py --gpus N --steps K --warmup W [--impl reference]`
prints ONE JSON line on rank 0.

Workload (BASELINE.json configs[1], "c2"): camera frames of 1920x1080 CV_8UC3 (row pitch 6144 B), 50 crops of
mixed size per frame (w ~ U{24..256}, h = 2w) -> 64x128 bilinear resize -> RGB2BGR -> *0.3 -> -sub -> /div ->
planar NCHW float.  One cvGS::executeOperations-equivalent call (= ONE kernel launch) per frame.
A *step* is one pass over `--frames` (default 32) distinct frames, each with its own source image, rect list and
output tensor, so that the working set (32 x 11.5 MB = 369 MB) is larger than the 126 MB L2 and every launch
reads its source from HBM and writes its tensor to HBM ("inputs larger than L2" rule).

  value      device-resident inputs: the C-ABI frame loop cvgs_b200_preproc_launch_sequence, timed with CUDA
             events on the launching stream; crops / s over all ranks (max time over ranks).
  e2e        the same loop through cvgs_b200_preproc_host_sequence: pinned HOST frames in, pinned HOST tensors
             out, H2D + kernel + D2H inside the timed region.
  roofline   algorithmic bytes of one launch (unique tapped source pixels x 3 B + 4 B x 3 x 64 x 128 x 50) /
             average launch duration inside the timed region, against MEASURED_PEAKS.json's HBM copy bandwidth.
  cpu_baseline / --impl reference
             the oracle port (oracle/liboracle.so, OpenMP on all host cores) on a bounded sample of the
             same frames.  The reference is a CUDA-only header library with no CPU implementation of this
             path (FKL's __host__ Interpolate has a typo, F10), so kind = "port".
  baselines  (rank 0, N=1) the reference's own fused GPU kernel instantiated from its headers
             (oracle/_ref/libfkref_50.so), a restated multi-kernel "OpenCV-CUDA-equivalent" chain
             (oracle/libchain.so) and OpenCV-CPU (cv2), same frames, same box.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

os.environ.setdefault("OMP_WAIT_POLICY", "passive")  # the CPU port's idle OpenMP threads must not spin

import numpy as np  # noqa: E402

DST = (64, 128)
CROPS_PER_FRAME = 50
FRAME = (1920, 1080)
PITCH = 6144
MUL, SUB, DIV = (0.3, 0.3, 0.3), (1.0, 4.0, 3.2), (3.2, 0.6, 11.8)
OPS = [("reorder", (2, 1, 0)), ("mul", MUL), ("sub", SUB), ("div", DIV)]


# --------------------------------------------------------------------------------------------------
# workload
# --------------------------------------------------------------------------------------------------
def make_frames(n_frames: int, seed: int, crops=CROPS_PER_FRAME):
    """n_frames x (image[H, pitch] uint8, rects[crops] (x, y, w, h)); SURVEY.md 8(d) C2 recipe."""
    rng = np.random.default_rng(seed)
    fw, fh = FRAME
    frames = []
    for _ in range(n_frames):
        img = rng.integers(0, 256, size=(fh, PITCH), dtype=np.uint8)
        rects = []
        for _ in range(crops):
            w = int(rng.integers(24, 257))
            h = min(2 * w, fh)
            rects.append((int(rng.integers(0, fw - w + 1)), int(rng.integers(0, fh - h + 1)), w, h))
        frames.append((img, rects))
    return frames


def algorithmic_bytes(rects, dst=DST, frame=FRAME) -> tuple[int, int]:
    """(bytes_in, bytes_out) of one launch, SURVEY.md 8(d): bytes_in = 3 x number of distinct source pixels
    that are a bilinear tap of at least one output pixel (index math of interpolation.cuh:57-92 with the
    scale of resize.cuh:100-114), shared between overlapping crops; bytes_out = 4 x 3 x W x H x planes."""
    fw, fh = frame
    dw, dh = dst
    mask = np.zeros((fh, fw), dtype=bool)
    for (x0, y0, w, h) in rects:
        fx = np.float32(1.0 / (float(dw) / float(w)))
        fy = np.float32(1.0 / (float(dh) / float(h)))
        sx = np.arange(dw, dtype=np.float32) * fx
        sy = np.arange(dh, dtype=np.float32) * fy
        x1 = np.floor(sx).astype(np.int64)
        y1 = np.floor(sy).astype(np.int64)
        cols = np.unique(np.concatenate([x1, np.minimum(x1 + 1, w - 1)])) + x0
        rows = np.unique(np.concatenate([y1, np.minimum(y1 + 1, h - 1)])) + y0
        mask[np.ix_(rows, cols)] = True
    return 3 * int(mask.sum()), 4 * 3 * dw * dh * len(rects)


# --------------------------------------------------------------------------------------------------
# clocks (sampled DURING the timed regions)
# --------------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    _REASONS = {0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
                0x80: "hw_power_brake_slowdown", 0x2: "applications_clocks_setting"}

    def __init__(self, device_index: int):
        super().__init__(daemon=True)
        self.samples, self.reasons, self.active, self._stop_evt = [], set(), False, threading.Event()
        self.sm_max = None
        self.ok = False
        # NVML queries take driver locks that kernel launches also need: sample sparsely so that the sampler does not
        # slow down the launch-bound loop it observes
        self.interval = float(os.environ.get("CVGS_BENCH_CLOCK_INTERVAL", "0.004"))
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            import torch
            try:
                uuid = str(torch.cuda.get_device_properties(device_index).uuid)
                if not uuid.startswith("GPU-"):
                    uuid = "GPU-" + uuid
                self.h = pynvml.nvmlDeviceGetHandleByUUID(uuid.encode())
            except Exception:
                self.h = pynvml.nvmlDeviceGetHandleByIndex(device_index)
            self.sm_max = int(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.ok = True
        except Exception:
            self.ok = False

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        while not self._stop_evt.is_set():
            if self.active:
                try:
                    self.samples.append(int(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                    try:
                        bits = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
                    except Exception:
                        bits = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                    for b, name in self._REASONS.items():
                        if bits & b:
                            self.reasons.add(name)
                except Exception:
                    pass
                time.sleep(self.interval)
            else:
                time.sleep(0.0005)

    def stop(self):
        self._stop_evt.set()

    def summary(self):
        if not self.ok or not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.sm_max, "reasons": sorted(self.reasons), "samples": 0}
        return {"sm_mhz": statistics.median(self.samples), "sm_max_mhz": self.sm_max,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


# --------------------------------------------------------------------------------------------------
# CPU arm: the oracle port on the host cores (cpu_baseline leg and --impl reference)
# --------------------------------------------------------------------------------------------------
_THREADS = None


def _best_thread_count(lib, crops, p) -> int:
    """Thread count that is actually fastest on this host (a cgroup CPU quota can make the full
    core count slower than a few threads); calibrated once on one frame."""
    global _THREADS
    if _THREADS is None:
        # the cores this process may run on -- not omp_get_max_threads(): torchrun exports OMP_NUM_THREADS=1, and the
        # num_threads clause of the port overrides that variable
        try:
            most = max(1, len(os.sched_getaffinity(0)))
        except AttributeError:
            most = max(1, os.cpu_count() or 1)
        cands = sorted({1, most} | {n for n in (2, 4, 8, 16, 32, 64, 128) if n < most})
        best = (float("inf"), 1)
        for n in cands:
            lib.oracle_preproc(crops, CROPS_PER_FRAME, CROPS_PER_FRAME, C.byref(p), n)  # spin up the pool
            t = time.perf_counter()
            lib.oracle_preproc(crops, CROPS_PER_FRAME, CROPS_PER_FRAME, C.byref(p), n)
            best = min(best, (time.perf_counter() - t, n))
        _THREADS = best[1]
    return _THREADS


def cpu_port_crops_per_s(frames, min_seconds: float, max_frames: int):
    """Times oracle_preproc (OpenMP, all cores) frame by frame; returns (crops/s, cores, frames done, seconds)."""
    from tests import util  # oracle loader: checker / CPU baseline only
    lib = util.oracle_lib()
    out = np.empty((CROPS_PER_FRAME, 3, DST[1], DST[0]), dtype=np.float32)
    p = util.make_pipeline(DST, OPS, out_ptr=out.ctypes.data)
    crop_sets = [util.host_crops(img, rects) for img, rects in frames]
    cores = _best_thread_count(lib, crop_sets[0], p)
    done, t0 = 0, time.perf_counter()
    while True:
        lib.oracle_preproc(crop_sets[done % len(frames)], CROPS_PER_FRAME, CROPS_PER_FRAME, C.byref(p), cores)
        done += 1
        dt = time.perf_counter() - t0
        if (dt >= min_seconds and done >= 4) or done >= max_frames:
            break
    return done * CROPS_PER_FRAME / dt, cores, done, dt


def opencv_cpu_crops_per_s(frames, min_seconds: float):
    """OpenCV-CPU chain (BASELINE.md baseline C).  Different resize semantics: throughput only."""
    try:
        import cv2
    except Exception:
        return None
    cv2.setNumThreads(os.cpu_count() or 1)
    done, t0 = 0, time.perf_counter()
    while True:
        img, rects = frames[done % len(frames)]
        view = img[:, :3 * FRAME[0]].reshape(FRAME[1], FRAME[0], 3)
        for (x, y, w, h) in rects:
            r = cv2.resize(view[y:y + h, x:x + w], DST, interpolation=cv2.INTER_LINEAR)
            f = r.astype(np.float32) * np.float32(0.3)
            f = cv2.subtract(f, SUB + (0.0,))
            f = cv2.divide(f, DIV + (1.0,))
            cv2.split(f)
        done += 1
        dt = time.perf_counter() - t0
        if dt >= min_seconds and done >= 2:
            break
    return {"value": done * CROPS_PER_FRAME / dt, "unit": "crops/s", "cores": os.cpu_count(),
            "threads": cv2.getNumThreads(), "sample": f"{done} frames x 50 crops", "version": cv2.__version__,
            "note": "cv2.resize uses half-pixel centres: throughput baseline, not a parity oracle"}


# --------------------------------------------------------------------------------------------------
# reference arm
# --------------------------------------------------------------------------------------------------
def run_reference_arm(args, rank: int, world: int):
    if rank != 0:
        return
    frames = make_frames(min(args.frames, 8), seed=2)
    per_step_frames = 4  # bounded sample of the step: 4 frames x 50 crops
    for _ in range(args.warmup):
        cpu_port_crops_per_s(frames, 0.0, per_step_frames)
    t0 = time.perf_counter()
    total = 0
    for _ in range(args.steps):
        _, cores, done, _ = cpu_port_crops_per_s(frames, 0.0, per_step_frames)
        total += done
    dt = time.perf_counter() - t0
    value = total * CROPS_PER_FRAME / dt
    line = {
        "impl": "reference", "metric": "crops_per_second", "value": value, "unit": "crops/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args.frames),
        "cpu_baseline": {"value": value, "unit": "crops/s", "cores": cores, "kind": "port",
                         "sample": f"each step = {per_step_frames} frames x 50 crops of the workload "
                                   f"(the GPU arm's step is {args.frames} frames)"},
        "e2e": {"value": value, "unit": "crops/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "the reference is a CUDA-only header library without a CPU implementation of this path; this arm "
                "times the CPU port of its algorithm (oracle/oracle.c, OpenMP, all host cores)",
    }
    print(json.dumps(line), flush=True)


def workload_config(n_frames: int):
    return {"workload": "c2: 50 crops/frame (mixed sizes, w~U{24..256}, h=2w) from 1920x1080 CV_8UC3 -> 64x128 "
                        "bilinear resize + RGB2BGR + mul/sub/div + NCHW split; one launch per frame",
            "frames_per_step": n_frames, "crops_per_step": n_frames * CROPS_PER_FRAME,
            "l2_policy": f"inputs larger than L2: {n_frames} rotating frame/tensor sets = "
                         f"{n_frames * (FRAME[1] * PITCH + CROPS_PER_FRAME * 3 * DST[0] * DST[1] * 4) / 1e6:.0f} MB",
            "fp_contract": "reference_fused", "interp_mode": "float", "sharding": "frames per GPU, no collective",
            "launch_api": "cvgs_b200_preproc_launch_sequence_ex -> one cvgs_b200_preproc_launch_ex (crops + parent frame) "
                          "per frame; consecutive independent frames may overlap (cvgs_b200_set_overlap(1)) and the "
                          "frame loop is driven by several host threads, one stream each",
            "host_threads": int(os.environ.get("CVGS_B200_SEQ_THREADS", "3"))}


# --------------------------------------------------------------------------------------------------
# GPU arm
# --------------------------------------------------------------------------------------------------
def gpu_baselines(frames, d_imgs, torch, min_seconds=0.5):
    """Reference fused kernel (its own headers, BATCH=50 instantiation) on the same device frames."""
    out = {}
    path = os.path.join(ROOT, "oracle", "_ref", "libfkref_50.so")
    if os.path.exists(path):
        lib = C.CDLL(path)
        fn = lib.fkref_preproc_50
        fn.restype = C.c_int
        fn.argtypes = [C.POINTER(C.c_void_p), C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int), C.c_int,
                       C.c_int, C.c_int, C.c_int, C.POINTER(C.c_float), C.c_int, C.POINTER(C.c_float),
                       C.POINTER(C.c_float), C.POINTER(C.c_float), C.c_void_p, C.c_void_p]
        f3 = lambda v: (C.c_float * 3)(*v)  # noqa: E731
        n = CROPS_PER_FRAME
        argsets = []
        outs = [torch.empty((n, 3, DST[1], DST[0]), dtype=torch.float32, device="cuda") for _ in frames]
        for (img, rects), d in zip(frames, d_imgs):
            base = d.data_ptr()
            ptrs = (C.c_void_p * n)(*[base + y * PITCH + 3 * x for (x, y, w, h) in rects])
            argsets.append((ptrs, (C.c_int * n)(*[r[2] for r in rects]), (C.c_int * n)(*[r[3] for r in rects]),
                            (C.c_int * n)(*[PITCH] * n)))
        bg, mul, sub, div = f3((0, 0, 0)), f3(MUL), f3(SUB), f3(DIV)
        s = torch.cuda.current_stream()
        nf = len(frames)
        seq = getattr(lib, "fkref_preproc_sequence_50", None)
        if seq is not None:  # native frame loop: no per-call ctypes overhead charged to the reference
            seq.restype = C.c_int
            P, I = C.POINTER, C.c_int
            seq.argtypes = [P(P(C.c_void_p)), P(P(I)), P(P(I)), P(P(I)), I, I, I, I, P(C.c_float), I, P(C.c_float),
                            P(C.c_float), P(C.c_float), P(C.c_void_p), I, I, C.c_void_p]
            a_ptrs = (P(C.c_void_p) * nf)(*[C.cast(a[0], P(C.c_void_p)) for a in argsets])
            a_ws = (P(I) * nf)(*[C.cast(a[1], P(I)) for a in argsets])
            a_hs = (P(I) * nf)(*[C.cast(a[2], P(I)) for a in argsets])
            a_ps = (P(I) * nf)(*[C.cast(a[3], P(I)) for a in argsets])
            a_outs = (C.c_void_p * nf)(*[o.data_ptr() for o in outs])

            def passes(k):
                rc = seq(a_ptrs, a_ws, a_hs, a_ps, n, DST[0], DST[1], 1, bg, 1, mul, sub, div, a_outs, nf, nf * k,
                         s.cuda_stream)
                assert rc == 0
            how = "launched from a native frame loop"
        else:
            def passes(k):
                for _ in range(k):
                    for (ptrs, ws, hs, ps), o in zip(argsets, outs):
                        rc = fn(ptrs, ws, hs, ps, n, DST[0], DST[1], 1, bg, 1, mul, sub, div, o.data_ptr(), s.cuda_stream)
                        assert rc == 0
            how = "launched from a ctypes loop"
        passes(20)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 50
        e0.record(s)
        passes(reps)
        e1.record(s)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        launches = reps * nf
        out["reference_fused_kernel_gpu"] = {
            "value": launches * n / (ms * 1e-3), "unit": "crops/s", "us_per_launch": ms * 1e3 / launches,
            "what": "fk::executeOperations(BatchRead<50>(Resize<INTER_LINEAR>), ColorConversion, Mul, Sub, Div, "
                    "TensorSplit) from /root/reference/fkl/include compiled for sm_100a (oracle/_ref/libfkref_50.so), "
                    "same device frames, " + how}
    out.update(chain_baseline(frames, d_imgs, torch))
    return out


def chain_baseline(frames, d_imgs, torch):
    """Baseline M: the restated multi-kernel OpenCV-CUDA-equivalent chain (oracle/chain_kernels.cu), 300 launches per
    50-crop frame, same device frames, native frame loop."""
    path = os.path.join(ROOT, "oracle", "libchain.so")
    if not os.path.exists(path):
        return {}
    lib = C.CDLL(path)
    P, I = C.POINTER, C.c_int
    lib.chain_workspace_bytes.restype = C.c_size_t
    lib.chain_workspace_bytes.argtypes = [I, I]
    seq = lib.chain_preproc_sequence
    seq.restype = I
    seq.argtypes = [P(P(C.c_void_p)), P(P(I)), P(P(I)), P(P(I)), I, I, I, I, P(C.c_float), P(C.c_float), P(C.c_float),
                    P(C.c_void_p), C.c_void_p, I, I, C.c_void_p]
    n, nf = CROPS_PER_FRAME, len(frames)
    keep = []
    for (img, rects), d in zip(frames, d_imgs):
        base = d.data_ptr()
        keep.append(((C.c_void_p * n)(*[base + y * PITCH + 3 * x for (x, y, w, h) in rects]),
                     (I * n)(*[r[2] for r in rects]), (I * n)(*[r[3] for r in rects]), (I * n)(*[PITCH] * n)))
    outs = [torch.empty((n, 3, DST[1], DST[0]), dtype=torch.float32, device="cuda") for _ in frames]
    ws = torch.empty(int(lib.chain_workspace_bytes(DST[0], DST[1])) + 512, dtype=torch.uint8, device="cuda")
    a_ptrs = (P(C.c_void_p) * nf)(*[C.cast(k[0], P(C.c_void_p)) for k in keep])
    a_ws = (P(I) * nf)(*[C.cast(k[1], P(I)) for k in keep])
    a_hs = (P(I) * nf)(*[C.cast(k[2], P(I)) for k in keep])
    a_ps = (P(I) * nf)(*[C.cast(k[3], P(I)) for k in keep])
    a_outs = (C.c_void_p * nf)(*[o.data_ptr() for o in outs])
    f3 = lambda v: (C.c_float * 3)(*v)  # noqa: E731
    s = torch.cuda.current_stream()

    def passes(k):
        rc = seq(a_ptrs, a_ws, a_hs, a_ps, n, DST[0], DST[1], 1, f3(MUL), f3(SUB), f3(DIV), a_outs, ws.data_ptr(), nf, nf * k,
                 s.cuda_stream)
        assert rc > 0
        return rc
    passes(1)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 3
    e0.record(s)
    launches = passes(reps)
    e1.record(s)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    return {"opencv_cuda_equivalent_chain_gpu": {
        "value": reps * nf * n / (ms * 1e-3), "unit": "crops/s", "us_per_frame": ms * 1e3 / (reps * nf),
        "launches_per_frame": launches // (reps * nf),
        "what": "restated multi-kernel chain per crop: resize(8UC3) -> convertTo(32F, alpha) -> cvtColor -> subtract -> "
                "divide -> split (oracle/chain_kernels.cu; real OpenCV-CUDA is not installable here), same device "
                "frames, native frame loop; equals the product's (SEPARATE, ROUND_U8) mode bit for bit"}}


def run_gpu_arm(args, rank: int, world: int, local_rank: int):
    import torch
    from cvgpuspeedup_b200 import _abi
    from tests import util

    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device (the product has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist_
        dist = dist_
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    lib = _abi.load()
    F, K, W = args.frames, args.steps, args.warmup

    frames = make_frames(F, seed=2 + 1000 * rank)
    alg = [algorithmic_bytes(r) for _, r in frames]
    bytes_in = sum(a for a, _ in alg) / F
    bytes_out = sum(b for _, b in alg) / F

    h_imgs = [torch.from_numpy(img).pin_memory() for img, _ in frames]
    d_imgs = [h.cuda() for h in h_imgs]
    d_outs = [torch.empty((CROPS_PER_FRAME, 3, DST[1], DST[0]), dtype=torch.float32, device="cuda") for _ in frames]
    h_outs = [torch.empty((CROPS_PER_FRAME, 3, DST[1], DST[0]), dtype=torch.float32).pin_memory() for _ in frames]

    # argument sets of the C-ABI frame loops
    crop_sets = [util.host_crops(img, rects, base_ptr=d.data_ptr()) for (img, rects), d in zip(frames, d_imgs)]
    pipes = [util.make_pipeline(DST, OPS, out_ptr=o.data_ptr()) for o in d_outs]
    parent_sets = [util.host_parents(img, FRAME[0], FRAME[1], CROPS_PER_FRAME, base_ptr=d.data_ptr())
                   for (img, _), d in zip(frames, d_imgs)]
    crops_pp = (C.POINTER(_abi.Crop) * F)(*[C.cast(c, C.POINTER(_abi.Crop)) for c in crop_sets])
    parents_pp = (C.POINTER(_abi.Parent) * F)(*[C.cast(c, C.POINTER(_abi.Parent)) for c in parent_sets])
    pipes_pp = (C.POINTER(_abi.Pipeline) * F)(*[C.pointer(p) for p in pipes])
    n_arr = (C.c_int32 * F)(*[CROPS_PER_FRAME] * F)
    rect_sets = [(_abi.Rect * CROPS_PER_FRAME)(*[_abi.Rect(*r) for r in rects]) for _, rects in frames]
    rects_pp = (C.POINTER(_abi.Rect) * F)(*[C.cast(r, C.POINTER(_abi.Rect)) for r in rect_sets])
    himg_pp = (C.c_void_p * F)(*[h.data_ptr() for h in h_imgs])
    hout_pp = (C.c_void_p * F)(*[h.data_ptr() for h in h_outs])
    host_pipe = util.make_pipeline(DST, OPS)
    hpipes_pp = (C.POINTER(_abi.Pipeline) * F)(*[C.pointer(host_pipe)] * F)

    stream = torch.cuda.Stream()
    sp = stream.cuda_stream

    # What the header shim emits for cvGS::executeOperations on GpuMat ROIs of a frame: the crops plus the frame they
    # were cut from (GpuMat::datastart / locateROI).  Consecutive frames are independent; the library is allowed to
    # prove that and overlap them (cvgs_b200_set_overlap, see include/cvgs_b200.h).
    lib.cvgs_b200_set_overlap(0 if args.no_overlap else 1)

    def device_steps(n):
        _abi.check(lib.cvgs_b200_preproc_launch_sequence_ex(crops_pp, parents_pp, n_arr, n_arr, pipes_pp, F, F * n, sp))

    def host_steps(n):
        _abi.check(lib.cvgs_b200_preproc_host_sequence(himg_pp, FRAME[0], FRAME[1], PITCH, rects_pp, n_arr, n_arr,
                                                       hpipes_pp, hout_pp, F, F * n, sp))

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)
    sampler.start()

    def timed(fn, n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        sampler.active = True
        l0 = lib.cvgs_b200_launch_count()
        e0.record(stream)
        fn(n)
        e1.record(stream)
        torch.cuda.synchronize()
        sampler.active = False
        launches = lib.cvgs_b200_launch_count() - l0
        ms = e0.elapsed_time(e1)
        if dist is not None:
            t = torch.tensor([ms], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        barrier()
        return ms, launches

    # ---- device-resident arm ----
    # W warm-up steps as asked, and at least ~100 ms of the same loop: a step is only 32 launches (~0.15 ms), and the
    # first thousands of launches after an idle period run slower (clock / power-state ramp)
    device_steps(max(W, 3, 640))
    ms_dev, launches = timed(device_steps, K)
    # parity spot check of what was just timed (frame 0 against the oracle) -- checker only
    want = util.run_oracle(frames[0][0], frames[0][1], DST, OPS)
    util.assert_bit_equal(d_outs[0].cpu().numpy(), want, "bench: frame 0 vs oracle")

    # ---- end-to-end arm (host buffers) ----
    host_steps(max(W, 3, 20))
    ms_e2e, launches_e2e = timed(host_steps, K)
    util.assert_bit_equal(h_outs[F - 1].numpy(), util.run_oracle(frames[F - 1][0], frames[F - 1][1], DST, OPS),
                          "bench: e2e last frame vs oracle")
    extra = None
    graph_extra = None
    if world == 1 and not args.no_baselines:
        extra = c3_extra(lib, torch, _abi, util, stream)
        try:  # the same 32-frame loop captured once into a CUDA graph and replayed: no host launch cost at all
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=stream):
                device_steps(1)
            for _ in range(20):
                g.replay()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            reps = 200
            e0.record()
            for _ in range(reps):
                g.replay()
            e1.record()
            torch.cuda.synchronize()
            us = e0.elapsed_time(e1) * 1e3 / (reps * F)
            util.assert_bit_equal(d_outs[1].cpu().numpy(), util.run_oracle(frames[1][0], frames[1][1], DST, OPS),
                                  "bench: graph replay frame 1 vs oracle")
            graph_extra = {"what": "the 32 launches of a step captured into one CUDA graph and replayed (crops fixed at "
                                   "capture time: an upper bound for pipelines whose crops change every frame)",
                           "us_per_launch": us, "crops_per_s": CROPS_PER_FRAME / (us * 1e-6),
                           "achieved_gbs": (bytes_in + bytes_out) / (us * 1e-6) / 1e9}
        except Exception as e:
            graph_extra = {"error": str(e)[:200]}
    sampler.stop()

    crops_total = world * F * K * CROPS_PER_FRAME
    value = crops_total / (ms_dev * 1e-3)
    e2e_value = crops_total / (ms_e2e * 1e-3)
    us_per_launch = ms_dev * 1e3 / (F * K)
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    achieved = (bytes_in + bytes_out) / (us_per_launch * 1e-6) / 1e9
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        try:
            traffic = json.load(open(tpath)).get("c2_dram_bytes_per_launch")
        except Exception:
            traffic = None

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    line = {
        "metric": "crops_per_second", "value": value, "unit": "crops/s", "n_gpus": world, "steps": K, "warmup": max(W, 3),
        "ms_per_step": ms_dev / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": workload_config(F),
        "e2e": {"value": e2e_value, "unit": "crops/s", "h2d_bytes_per_step": int(h2d_bytes(frames)),
                "d2h_bytes_per_step": int(bytes_out * F), "ms_per_step": ms_e2e / K,
                "api": "cvgs_b200_preproc_host_sequence (pinned host frames -> pinned host tensors, 3 frames in "
                       "flight: upload / kernel / download overlap)"},
        "gpu_launches": int(launches),
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic, "kernel": "preproc kernel (one launch per frame)", "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": bytes_in + bytes_out, "bytes_in": bytes_in, "bytes_out": bytes_out,
                     "us_per_launch": us_per_launch,
                     "note": "us_per_launch = timed region / launches (frames of different host threads overlap on the "
                             "GPU, so this is the effective per-launch time); 50 crops = 7.5 MB per launch, one host "
                             "thread alone is launch-bound at ~3.9 us (SURVEY F6); see c3 in 'extra' for the same "
                             "kernel on a 256-crop batch"},
        "clocks": sampler.summary(),
    }
    if extra:
        extra["frac_of_peak"] = extra["achieved_gbs"] / peak
        line["extra"] = {"c3": extra}
        if graph_extra:
            if "achieved_gbs" in graph_extra:
                graph_extra["frac_of_peak"] = graph_extra["achieved_gbs"] / peak
            line["extra"]["c2_cuda_graph"] = graph_extra
        try:
            line["extra"]["c2_single_launch"] = c2_latency_extra(lib, torch, _abi, crops_pp, parents_pp, n_arr, pipes_pp, F, stream)
        except Exception as e:
            line["extra"]["c2_single_launch"] = {"error": str(e)[:200]}
        try:
            c4 = c4_extra(torch, util, stream)
            c4["frac_of_peak"] = c4["achieved_gbs"] / peak
            line["extra"]["c4"] = c4
        except Exception as e:  # the extra lines never take the headline down with them
            line["extra"]["c4"] = {"error": str(e)[:200]}
    if world == 1:
        cps, cores, done, dt = cpu_port_crops_per_s(frames, args.cpu_seconds, 10 ** 9)
        line["cpu_baseline"] = {"value": cps, "unit": "crops/s", "cores": cores, "kind": "port",
                                "sample": f"{done} frames x 50 crops of the same workload in {dt:.1f} s "
                                          "(oracle/oracle.c, OpenMP)"}
        if not args.no_baselines:
            b = gpu_baselines(frames, d_imgs, torch)
            ocv = opencv_cpu_crops_per_s(frames, min(args.cpu_seconds, 3.0))
            if ocv:
                b["opencv_cpu"] = ocv
            line["baselines"] = b
    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


def c3_extra(lib, torch, _abi, util, stream, reps=20):
    """BASELINE configs[2]: 256 crops (224..896 px) of a 3840x2160 frame -> 224x224, BGR2RGB, ImageNet mean/std,
    NCHW.  154 MB of output per launch, 2 rotating (frame, tensor) sets > L2: the configuration on which the HBM
    roofline fraction describes the kernel rather than launch latency."""
    sets = []
    for k in range(2):
        w = util.workload_c3(seed=3 + k)
        d_img = torch.from_numpy(w.image).cuda()
        d_out = torch.empty((256, 3, 224, 224), dtype=torch.float32, device="cuda")
        crops = util.host_crops(w.image, w.rects, base_ptr=d_img.data_ptr())
        pipe = util.make_pipeline(w.dsize, w.ops, out_ptr=d_out.data_ptr())
        par = util.host_parents(w.image, w.width, w.height, len(w.rects), base_ptr=d_img.data_ptr())
        sets.append((w, d_img, d_out, crops, pipe, par))
    n = len(sets)
    crops_pp = (C.POINTER(_abi.Crop) * n)(*[C.cast(s[3], C.POINTER(_abi.Crop)) for s in sets])
    pipes_pp = (C.POINTER(_abi.Pipeline) * n)(*[C.pointer(s[4]) for s in sets])
    par_pp = (C.POINTER(_abi.Parent) * n)(*[C.cast(s[5], C.POINTER(_abi.Parent)) for s in sets])
    n_arr = (C.c_int32 * n)(*[256] * n)
    sp = stream.cuda_stream
    _abi.check(lib.cvgs_b200_preproc_launch_sequence_ex(crops_pp, par_pp, n_arr, n_arr, pipes_pp, n, 4, sp))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record(stream)
    _abi.check(lib.cvgs_b200_preproc_launch_sequence_ex(crops_pp, par_pp, n_arr, n_arr, pipes_pp, n, reps, sp))
    e1.record(stream)
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / reps
    w0 = sets[0][0]
    idx = [0, 100, 255]
    want = util.run_oracle(w0.image, [w0.rects[i] for i in idx], w0.dsize, w0.ops)
    util.assert_bit_equal(sets[0][2][idx].cpu().numpy(), want, "bench c3: spot check vs oracle")
    b_in, b_out = algorithmic_bytes(w0.rects, dst=(224, 224), frame=(3840, 2160))
    gbs = (b_in + b_out) / (us * 1e-6) / 1e9
    ref_us = c3_reference_us(sets, torch, stream)
    return {"reference_fused_kernel_us_per_256_crops": ref_us,
            "reference_note": "fk::executeOperations with BATCH=128 (a template parameter capped at 255, SURVEY F7): two "
                              "launches per 256 crops, same frames, oracle/_ref/libfkref_128.so",
            "workload": "c3: 256 crops (224..896 px) from 3840x2160 -> 224x224 + BGR2RGB + mean/std + NCHW, one launch",
            "us_per_launch": us, "crops_per_s": 256 / (us * 1e-6), "algorithmic_bytes_per_launch": b_in + b_out,
            "bytes_in": b_in, "bytes_out": b_out, "achieved_gbs": gbs}


def c4_extra(torch, util, stream, depth=16, reps=20):
    """BASELINE configs[3]: CircularTensor depth 16 of 1920x1080 planes, 1080p CV_8UC3 frames (no resize), BGR2RGB +
    mean/std on the new frame; one kernel per update shifts the other 15 planes and processes the new one."""
    import cvgpuspeedup_b200 as cvGS
    W, H = 1920, 1080
    rng = np.random.default_rng(4)
    frames = [torch.from_numpy(util.make_image(rng, W, H, pitch=6144)).cuda() for _ in range(4)]
    mats = [cvGS.GpuMat(f.data_ptr(), W, H, 6144, owner=f) for f in frames]
    ops = [cvGS.cvtColor(cvGS.COLOR_BGR2RGB), cvGS.multiply((1 / 255.0,) * 3), cvGS.subtract((0.485, 0.456, 0.406)),
           cvGS.divide((0.229, 0.224, 0.225))]
    ct = cvGS.CircularTensor(W, H, depth, cvGS.CT_NEWEST_FIRST, cvGS.CT_STANDARD)
    try:
        for i in range(depth + 2):
            ct.update(stream, mats[i % 4], *ops)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for i in range(reps):
            ct.update(stream, mats[i % 4], *ops)
        e1.record(stream)
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / reps
    finally:
        ct.close()
    plane = 12 * W * H
    alg = 3 * W * H + (depth - 1) * plane + depth * plane
    return {"workload": f"c4: CircularTensor depth {depth}, {W}x{H} planes, 1080p frames, one kernel per update",
            "us_per_update": us, "updates_per_s": 1e6 / us, "algorithmic_bytes_per_update": alg,
            "achieved_gbs": alg / (us * 1e-6) / 1e9}


def c2_latency_extra(lib, torch, _abi, crops_pp, parents_pp, n_arr, pipes_pp, n_sets, stream, reps=200):
    """the latency of ONE 50-crop launch (stream idle before and after: event, launch, event, synchronise),
    plain stream order, min / median over `reps` launches on rotating frames; and the back-to-back time of the same
    launches from one host thread (the frame loop of the headline uses three)."""
    prev = lib.cvgs_b200_set_overlap(0)
    try:
        sp = stream.cuda_stream
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        lat = []
        for i in range(reps + 20):
            s = i % n_sets
            torch.cuda.synchronize()
            e0.record(stream)
            _abi.check(lib.cvgs_b200_preproc_launch_ex(crops_pp[s], parents_pp[s], n_arr[s], n_arr[s], pipes_pp[s], sp))
            e1.record(stream)
            torch.cuda.synchronize()
            if i >= 20:
                lat.append(e0.elapsed_time(e1) * 1e3)
        lat.sort()
        torch.cuda.synchronize()
        e0.record(stream)
        _abi.check(lib.cvgs_b200_preproc_launch_sequence_ex(crops_pp, parents_pp, n_arr, n_arr, pipes_pp, n_sets, 20 * n_sets, sp))
        e1.record(stream)
        torch.cuda.synchronize()
        serial = e0.elapsed_time(e1) * 1e3 / (20 * n_sets)
    finally:
        lib.cvgs_b200_set_overlap(prev)
    return {"what": "one isolated 50-crop launch between two events on an idle stream (includes the event overhead), plain "
                    "stream order; and the same launches back to back from one host thread without overlap",
            "latency_us_min": lat[0], "latency_us_median": lat[len(lat) // 2], "launches": len(lat),
            "one_thread_stream_order_us_per_launch": serial}


def c3_reference_us(sets, torch, stream, reps=10):
    """The reference's own fused kernel on the c3 sets: 2 launches of 128 crops (its batch is a template parameter)."""
    path = os.path.join(ROOT, "oracle", "_ref", "libfkref_128.so")
    if not os.path.exists(path):
        return None
    lib = C.CDLL(path)
    fn = lib.fkref_preproc_128
    fn.restype = C.c_int
    fn.argtypes = [C.POINTER(C.c_void_p), C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int), C.c_int,
                   C.c_int, C.c_int, C.c_int, C.POINTER(C.c_float), C.c_int, C.POINTER(C.c_float),
                   C.POINTER(C.c_float), C.POINTER(C.c_float), C.c_void_p, C.c_void_p]
    f3 = lambda v: (C.c_float * 3)(*v)  # noqa: E731
    calls = []
    for (w, d_img, d_out, _c, _p, _par) in sets:
        ref_out = torch.empty_like(d_out)
        for half in range(2):
            rects = w.rects[128 * half:128 * (half + 1)]
            ptrs = (C.c_void_p * 128)(*[d_img.data_ptr() + y * w.pitch + 3 * x for (x, y, _, _) in rects])
            calls.append((ptrs, (C.c_int * 128)(*[r[2] for r in rects]), (C.c_int * 128)(*[r[3] for r in rects]),
                          (C.c_int * 128)(*[w.pitch] * 128), ref_out[128 * half:].data_ptr(), ref_out))
    mul, sub, div = (1 / 255.0,) * 3, (0.485, 0.456, 0.406), (0.229, 0.224, 0.225)

    def one_pass():
        for ptrs, ws, hs, ps, optr, _keep in calls:
            rc = fn(ptrs, ws, hs, ps, 128, 224, 224, 1, f3((0, 0, 0)), 1, f3(mul), f3(sub), f3(div), optr, stream.cuda_stream)
            assert rc == 0
    one_pass()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(reps):
        one_pass()
    e1.record(stream)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / (reps * len(sets))


def h2d_bytes(frames) -> int:
    """Bytes cvgs_b200_preproc_host uploads per step: the rows [min y, max y+h) of each frame, 3*W bytes each."""
    total = 0
    for _, rects in frames:
        lo = min(r[1] for r in rects)
        hi = max(r[1] + r[3] for r in rects)
        total += (hi - lo) * 3 * FRAME[0]
    return total


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--frames", type=int, default=32, help="distinct frame/tensor sets per step (> L2 in total)")
    ap.add_argument("--cpu-seconds", type=float, default=5.0, help="wall-clock budget of the CPU baseline sample")
    ap.add_argument("--no-baselines", action="store_true")
    ap.add_argument("--no-overlap", action="store_true", help="plain stream order between consecutive launches")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference_arm(args, rank, world)
        return
    if world != args.gpus and world == 1 and args.gpus > 1:
        raise SystemExit("launch with torchrun --nproc-per-node N for --gpus N")
    run_gpu_arm(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
