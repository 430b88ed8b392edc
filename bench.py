#!/usr/bin/env python
"""bench.py -- crops/s of the fused crop -> resize -> normalise -> split path on B200.

Contract (see the task statement): `python bench.py --gpus N --steps K --warmup W [--impl reference]`
prints ONE JSON line on rank 0.

Workload (BASELINE.json configs[1], "c2"): camera frames of 1920x1080 CV_8UC3 (row pitch 6144 B), 50 crops of
mixed size per frame (w ~ U{24..256}, h = 2w) -> 64x128 bilinear resize -> RGB2BGR -> *0.3 -> -sub -> /div ->
planar NCHW float, every frame with its own source image, rect list and output tensor.
A *step* is one pass over `--frames` (default 32) distinct frames, so that the working set (32 x 11.5 MB = 369 MB)
is larger than the 126 MB L2 and every frame is read from HBM and written to HBM ("inputs larger than L2" rule).

  value      device-resident inputs: the C-ABI frame loop cvgs_b200_preproc_launch_sequence_ex (one host thread;
             consecutive independent frames share kernel launches), timed with CUDA events on the launching stream;
             crops / s over all ranks (max time over ranks).  The K-step region is repeated until >= 100 ms have
             been timed; value is the median repetition, the fastest one is reported beside it.
  e2e        the same frames through cvgs_b200_preproc_host_sequence: pinned HOST frames in, pinned HOST tensors
             out, H2D + kernel + D2H inside the timed region.
  roofline   algorithmic bytes of one frame (unique tapped source pixels x 3 B + 4 B x 3 x 64 x 128 x 50) /
             device time per frame inside the timed region, against MEASURED_PEAKS.json's HBM copy bandwidth.
  cpu_baseline / --impl reference
             the oracle port (oracle/liboracle.so, OpenMP on all host cores) on a bounded sample of the
             same frames.  The reference is a CUDA-only header library with no CPU implementation of this
             path (FKL's __host__ Interpolate has a typo, F10), so kind = "port".
  baselines  (rank 0, N=1) the reference's own fused GPU kernel instantiated from its headers
             (oracle/_ref/libfkref_50.so; one launch thread and three), a restated multi-kernel
             "OpenCV-CUDA-equivalent" chain (oracle/libchain.so) and OpenCV-CPU (cv2), same frames, same box.
  extra      c2 launched per frame (three threads / one thread / CUDA graph / isolated launch), c3 (256 crops of a
             4K frame -> 224x224), c4 (CircularTensor depth 16), c5 (8192 crops sharded over the ranks: kernel only,
             kernel + in-place NCCL all-gather, and the gather fused into the kernel's stores over peer memory).
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
OPS_C3 = [("reorder", (2, 1, 0)), ("mul", (1 / 255.0,) * 3), ("sub", (0.485, 0.456, 0.406)), ("div", (0.229, 0.224, 0.225))]


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


def make_c3(seed: int, n: int = 256, frame=(3840, 2160), lo=224, hi=896):
    """SURVEY.md 8(d) C3 / C5 recipe: n rects (w, h ~ U{224..896}, independent) of one 4K frame."""
    rng = np.random.default_rng(seed)
    fw, fh = frame
    img = rng.integers(0, 256, size=(fh, 3 * fw), dtype=np.uint8)
    rects = []
    for _ in range(n):
        w, h = int(rng.integers(lo, hi + 1)), int(rng.integers(lo, hi + 1))
        rects.append((int(rng.integers(0, fw - w + 1)), int(rng.integers(0, fh - h + 1)), w, h))
    return img, rects


def tap_mask(rects, dst, frame):
    """Boolean mask of the source pixels that are a bilinear tap of at least one output pixel (index math of
    interpolation.cuh:57-92 with the scale of resize.cuh:100-114), shared between overlapping crops."""
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
    return mask


def algorithmic_bytes(rects, dst=DST, frame=FRAME) -> tuple[int, int]:
    """(bytes_in, bytes_out) of one launch, SURVEY.md 8(d): bytes_in = 3 x number of distinct tapped source pixels;
    bytes_out = 4 x 3 x W x H x planes."""
    return 3 * int(tap_mask(rects, dst, frame).sum()), 4 * 3 * dst[0] * dst[1] * len(rects)


def sector_bytes(rects, dst=DST, frame=FRAME, pitch=PITCH) -> int:
    """Secondary diagnostic of SURVEY.md 8(d): the tapped pixels at 32-byte sector granularity -- what any
    implementation must at least move from DRAM for the source side."""
    mask = tap_mask(rects, dst, frame)
    byte_mask = np.repeat(mask, 3, axis=1)
    pad = (-byte_mask.shape[1]) % 32
    if pad:
        byte_mask = np.pad(byte_mask, ((0, 0), (0, pad)))
    return 32 * int(byte_mask.reshape(byte_mask.shape[0], -1, 32).any(axis=2).sum())


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
        # slow down the loop it observes
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
    from cvgpuspeedup_b200 import marshal
    lib = util.oracle_lib()
    out = np.empty((CROPS_PER_FRAME, 3, DST[1], DST[0]), dtype=np.float32)
    p = marshal.make_pipeline(DST, OPS, out_ptr=out.ctypes.data)
    crop_sets = [marshal.crop_array(img.ctypes.data, img.shape[1], rects) for img, rects in frames]
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
    frames = make_frames(args.frames, seed=2)
    per_step_frames = args.frames  # the GPU arm's step: one pass over the frames
    for _ in range(min(args.warmup, 3)):
        cpu_port_crops_per_s(frames, 0.0, per_step_frames)
    t0 = time.perf_counter()
    total = 0
    steps = 0
    for _ in range(args.steps):
        _, cores, done, _ = cpu_port_crops_per_s(frames, 0.0, per_step_frames)
        total += done
        steps += 1
        if time.perf_counter() - t0 > 120.0:  # bounded: a slow host still ends within minutes
            break
    dt = time.perf_counter() - t0
    value = total * CROPS_PER_FRAME / dt
    line = {
        "impl": "reference", "metric": "crops_per_second", "value": value, "unit": "crops/s", "n_gpus": args.gpus,
        "steps": steps, "warmup": min(args.warmup, 3), "ms_per_step": dt / steps * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args.frames),
        "cpu_baseline": {"value": value, "unit": "crops/s", "cores": cores, "kind": "port",
                         "sample": f"each step = {per_step_frames} frames x 50 crops of the workload, {steps} steps"},
        "e2e": {"value": value, "unit": "crops/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "the reference is a CUDA-only header library without a CPU implementation of this path; this arm "
                "times the CPU port of its algorithm (oracle/oracle.c, OpenMP, all host cores)",
    }
    print(json.dumps(line), flush=True)


def workload_config(n_frames: int):
    return {"workload": "c2: 50 crops/frame (mixed sizes, w~U{24..256}, h=2w) from 1920x1080 CV_8UC3 -> 64x128 "
                        "bilinear resize + RGB2BGR + mul/sub/div + NCHW split; every frame has its own image, rects "
                        "and output tensor",
            "frames_per_step": n_frames, "crops_per_step": n_frames * CROPS_PER_FRAME,
            "l2_policy": f"inputs larger than L2: {n_frames} rotating frame/tensor sets = "
                         f"{n_frames * (FRAME[1] * PITCH + CROPS_PER_FRAME * 3 * DST[0] * DST[1] * 4) / 1e6:.0f} MB",
            "fp_contract": "reference_fused", "interp_mode": "float", "sharding": "frames per GPU, no collective",
            "launch_api": "cvgs_b200_preproc_launch_sequence_ex (crops + parent frame per frame) with "
                          "cvgs_b200_set_overlap(1): the library proves the frames independent and lets consecutive "
                          "frames share kernel launches (up to 928 crops / 32 frames per launch)",
            "host_threads": 1}


# --------------------------------------------------------------------------------------------------
# GPU arm
# --------------------------------------------------------------------------------------------------
class Frames:
    """The bench workload on one device: frames, tensors and the C-ABI argument sets."""

    def __init__(self, torch, frames):
        from cvgpuspeedup_b200 import marshal
        self.frames = frames
        self.F = len(frames)
        self.h_imgs = [torch.from_numpy(img).pin_memory() for img, _ in frames]
        self.d_imgs = [h.cuda() for h in self.h_imgs]
        shape = (CROPS_PER_FRAME, 3, DST[1], DST[0])
        self.d_outs = [torch.empty(shape, dtype=torch.float32, device="cuda") for _ in frames]
        self.sets = marshal.FrameSets([(d.data_ptr(), PITCH, FRAME[0], FRAME[1], rects)
                                       for (_, rects), d in zip(frames, self.d_imgs)],
                                      [o.data_ptr() for o in self.d_outs], DST, OPS)

    def device_steps(self, lib, n, stream_ptr):
        self.sets.launch_sequence(lib, self.F * n, stream_ptr)


def timed_repeats(torch, dist, stream, fn, k, sampler, min_total_ms=100.0, min_reps=5, max_reps=400):
    """fn(k) = k steps.  Repeats the K-step region (barrier + synchronise on both sides, CUDA events on the launching
    stream, max over ranks) until min_total_ms have been timed.  Returns the per-repetition times in ms."""
    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    times = []
    while len(times) < max_reps and (len(times) < min_reps or sum(times) < min_total_ms):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        if sampler is not None:
            sampler.active = True
        e0.record(stream)
        fn(k)
        e1.record(stream)
        torch.cuda.synchronize()
        if sampler is not None:
            sampler.active = False
        ms = e0.elapsed_time(e1)
        if dist is not None:
            t = torch.tensor([ms], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        times.append(ms)
        barrier()
    return times


def run_gpu_arm(args, rank: int, world: int, local_rank: int):
    import torch
    from cvgpuspeedup_b200 import _abi, marshal
    from tests import util  # oracle: checker and CPU legs only

    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device (the product has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist_
        dist = dist_
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    lib = _abi.load()
    F, K, W = args.frames, args.steps, max(args.warmup, 3)

    frames = make_frames(F, seed=2 + 1000 * rank)
    alg = [algorithmic_bytes(r) for _, r in frames]
    bytes_in = sum(a for a, _ in alg) / F
    bytes_out = sum(b for _, b in alg) / F
    bytes_in_sector = sum(sector_bytes(r) for _, r in frames) / F

    wl = Frames(torch, frames)
    h_outs = [torch.empty((CROPS_PER_FRAME, 3, DST[1], DST[0]), dtype=torch.float32).pin_memory() for _ in frames]
    n_arr = wl.sets.n_arr
    rect_sets = [(_abi.Rect * CROPS_PER_FRAME)(*[_abi.Rect(*r) for r in rects]) for _, rects in frames]
    rects_pp = (C.POINTER(_abi.Rect) * F)(*[C.cast(r, C.POINTER(_abi.Rect)) for r in rect_sets])
    himg_pp = (C.c_void_p * F)(*[h.data_ptr() for h in wl.h_imgs])
    hout_pp = (C.c_void_p * F)(*[h.data_ptr() for h in h_outs])
    host_pipe = marshal.make_pipeline(DST, OPS)
    hpipes_pp = (C.POINTER(_abi.Pipeline) * F)(*[C.pointer(host_pipe)] * F)

    stream = torch.cuda.Stream()
    sp = stream.cuda_stream

    # What the header shim emits for cvGS::executeOperations on GpuMat ROIs of a frame: the crops plus the frame they
    # were cut from (GpuMat::datastart / locateROI).  Consecutive frames are independent; the library is allowed to
    # prove that, overlap them and let them share launches (cvgs_b200_set_overlap, see include/cvgs_b200.h).
    lib.cvgs_b200_set_overlap(0 if args.no_overlap else 1)

    def device_steps(n):
        wl.device_steps(lib, n, sp)

    def host_steps(n):
        _abi.check(lib.cvgs_b200_preproc_host_sequence(himg_pp, FRAME[0], FRAME[1], PITCH, rects_pp, n_arr, n_arr,
                                                       hpipes_pp, hout_pp, F, F * n, sp))

    sampler = ClockSampler(local_rank)
    sampler.start()

    # ---- device-resident arm ----
    # W warm-up steps as asked, then the same loop for >= 100 ms of wall clock: a step is ~60 us of device time, and
    # the first launches after an idle period run slower (clock / power-state ramp).  The number of steps actually
    # run before timing is what the line reports as "warmup".
    warm_steps = 0
    device_steps(W)
    warm_steps += W
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    while time.perf_counter() - t0 < 0.1:
        device_steps(20)
        torch.cuda.synchronize()
        warm_steps += 20
    l0 = lib.cvgs_b200_launch_count()
    times = timed_repeats(torch, dist, stream, device_steps, K, sampler)
    launches_per_rep = (lib.cvgs_b200_launch_count() - l0) // len(times)
    ms_dev = statistics.median(times)
    # parity of what was just timed (first and last frame against the oracle) -- checker only
    for f in (0, F - 1):
        util.assert_bit_equal(wl.d_outs[f].cpu().numpy(), util.run_oracle(frames[f][0], frames[f][1], DST, OPS),
                              f"bench: frame {f} vs oracle")

    # ---- end-to-end arm (host buffers) ----
    host_steps(W)
    b_h2d, b_d2h = C.c_uint64(), C.c_uint64()
    lib.cvgs_b200_debug_host_bytes(None, None, 1)
    host_steps(1)
    lib.cvgs_b200_debug_host_bytes(C.byref(b_h2d), C.byref(b_d2h), 1)  # what one step actually moves, counted by the library
    times_e2e = timed_repeats(torch, dist, stream, host_steps, K, sampler, min_total_ms=300.0, min_reps=3, max_reps=20)
    ms_e2e = statistics.median(times_e2e)
    util.assert_bit_equal(h_outs[F - 1].numpy(), util.run_oracle(frames[F - 1][0], frames[F - 1][1], DST, OPS),
                          "bench: e2e last frame vs oracle")
    sampler.stop()

    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"

    extra = {}
    if not args.no_extras:
        if world == 1:
            for name, fn in (("c2_per_frame_launches", lambda: c2_per_frame_extra(lib, torch, wl, stream, bytes_in + bytes_out)),
                             ("c2_cuda_graph", lambda: c2_graph_extra(lib, torch, util, wl, frames, stream, bytes_in + bytes_out)),
                             ("c2_single_launch", lambda: c2_latency_extra(lib, torch, _abi, wl, stream)),
                             ("c3", lambda: c3_extra(lib, torch, _abi, marshal, util, stream)),
                             ("c4", lambda: c4_extra(torch, util, stream))):
                try:  # the extra lines never take the headline down with them
                    extra[name] = fn()
                except Exception as e:  # noqa: BLE001
                    extra[name] = {"error": f"{type(e).__name__}: {e}"[:300]}
        try:
            extra["c5"] = c5_extra(lib, torch, dist, _abi, marshal, util, rank, world, args.c5_crops)
        except Exception as e:  # noqa: BLE001
            extra["c5"] = {"error": f"{type(e).__name__}: {e}"[:300]}
        for v in extra.values():
            if isinstance(v, dict) and "achieved_gbs" in v:
                v["frac_of_peak"] = v["achieved_gbs"] / peak

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    crops_step = world * F * CROPS_PER_FRAME
    value = crops_step * K / (ms_dev * 1e-3)
    e2e_value = crops_step * K / (ms_e2e * 1e-3)
    us_per_frame = ms_dev * 1e3 / (F * K)
    achieved = (bytes_in + bytes_out) / (us_per_frame * 1e-6) / 1e9
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        try:
            traffic = json.load(open(tpath)).get("c2_dram_bytes_per_frame")
        except Exception:
            traffic = None

    line = {
        "metric": "crops_per_second", "value": value, "unit": "crops/s", "n_gpus": world, "steps": K, "warmup": warm_steps,
        "ms_per_step": ms_dev / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": workload_config(F),
        "timing": {"repetitions": len(times), "timed_ms_total": sum(times), "ms_per_step_median": ms_dev / K,
                   "ms_per_step_min": min(times) / K, "ms_per_step_max": max(times) / K,
                   "value_best_repetition": crops_step * K / (min(times) * 1e-3), "warmup_requested": args.warmup,
                   "note": "each repetition = K steps between barrier + synchronise, CUDA events on the launching "
                           "stream, max over ranks; value / ms_per_step are the median repetition"},
        "e2e": {"value": e2e_value, "unit": "crops/s", "h2d_bytes_per_step": int(b_h2d.value),
                "d2h_bytes_per_step": int(b_d2h.value), "ms_per_step": ms_e2e / K, "repetitions": len(times_e2e),
                "value_best_repetition": crops_step * K / (min(times_e2e) * 1e-3),
                "h2d_bytes_whole_rows": int(h2d_bytes(frames)),
                "api": "cvgs_b200_preproc_host_sequence (pinned host frames -> pinned host tensors, 3 frames in "
                       "flight: upload / kernel / download overlap)"},
        "gpu_launches": int(launches_per_rep),
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic, "kernel": "preproc_tma_kernel (shared by up to 18 frames of 50 crops)",
                     "peak_source": peak_src, "unit_of_work": "one 50-crop frame",
                     "algorithmic_bytes_per_frame": bytes_in + bytes_out, "bytes_in": bytes_in, "bytes_out": bytes_out,
                     "us_per_frame": us_per_frame, "frames_per_launch": F * K / max(1, launches_per_rep),
                     "bytes_in_at_sector_granularity": bytes_in_sector,
                     "frac_at_sector_granularity": (bytes_in_sector + bytes_out) / (us_per_frame * 1e-6) / 1e9 / peak,
                     "note": "achieved = algorithmic bytes of a frame / device time per frame in the timed region "
                             "(launches are back to back on one stream and overlap at their edges, so this is the "
                             "kernel's sustained rate); the algorithmic source bytes count tapped PIXELS (3 B each) -- "
                             "down-scales beyond 2x skip pixels that share 32-byte sectors with tapped ones, which is "
                             "what the sector-granularity figure adds"},
        "clocks": sampler.summary(),
    }
    if extra:
        line["extra"] = extra
    if world == 1:
        cps, cores, done, dt = cpu_port_crops_per_s(frames, args.cpu_seconds, 10 ** 9)
        line["cpu_baseline"] = {"value": cps, "unit": "crops/s", "cores": cores, "kind": "port",
                                "sample": f"{done} frames x 50 crops of the same workload in {dt:.1f} s "
                                          "(oracle/oracle.c, OpenMP)"}
        if not args.no_baselines:
            b = gpu_baselines(frames, wl.d_imgs, torch)
            ocv = opencv_cpu_crops_per_s(frames, min(args.cpu_seconds, 3.0))
            if ocv:
                b["opencv_cpu"] = ocv
            line["baselines"] = b
    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


def _event_us(torch, stream, fn, units):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record(stream)
    fn()
    e1.record(stream)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / units


def c2_per_frame_extra(lib, torch, wl, stream, bytes_frame):
    """The same frames with ONE launch per frame (cvgs_b200_set_coalesce(0)): three launch threads / streams (round 1's
    headline), and one thread in plain stream order."""
    sp = stream.cuda_stream
    out = {"what": "one kernel launch per 50-crop frame instead of shared launches"}
    prev = lib.cvgs_b200_set_coalesce(0)
    try:
        wl.device_steps(lib, 100, sp)
        us = _event_us(torch, stream, lambda: wl.device_steps(lib, 300, sp), 300 * wl.F)
        out["three_threads_overlap"] = {"us_per_frame": us, "crops_per_s": CROPS_PER_FRAME / (us * 1e-6),
                                        "achieved_gbs": bytes_frame / (us * 1e-6) / 1e9}
        prev_o = lib.cvgs_b200_set_overlap(0)
        try:
            wl.device_steps(lib, 20, sp)
            us = _event_us(torch, stream, lambda: wl.device_steps(lib, 100, sp), 100 * wl.F)
        finally:
            lib.cvgs_b200_set_overlap(prev_o)
        out["one_thread_stream_order"] = {"us_per_frame": us, "crops_per_s": CROPS_PER_FRAME / (us * 1e-6),
                                          "achieved_gbs": bytes_frame / (us * 1e-6) / 1e9}
    finally:
        lib.cvgs_b200_set_coalesce(prev)
    return out


def c2_graph_extra(lib, torch, util, wl, frames, stream, bytes_frame):
    """The launches of a step captured into one CUDA graph and replayed: no host launch cost at all."""
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=stream):
        wl.device_steps(lib, 1, torch.cuda.current_stream().cuda_stream)
    for _ in range(20):
        g.replay()
    reps = 200

    def replay():
        for _ in range(reps):
            g.replay()
    us = _event_us(torch, torch.cuda.current_stream(), replay, reps * wl.F)
    util.assert_bit_equal(wl.d_outs[1].cpu().numpy(), util.run_oracle(frames[1][0], frames[1][1], DST, OPS),
                          "bench: graph replay frame 1 vs oracle")
    return {"what": "the per-frame launches of a step captured into one CUDA graph and replayed (crops fixed at capture "
                    "time: for pipelines whose crops do not change; capture uses one launch per frame)",
            "us_per_frame": us, "crops_per_s": CROPS_PER_FRAME / (us * 1e-6), "achieved_gbs": bytes_frame / (us * 1e-6) / 1e9}


def c2_latency_extra(lib, torch, _abi, wl, stream, reps=200):
    """the latency of ONE 50-crop launch (stream idle before and after: event, launch, event, synchronise),
    plain stream order, min / median over `reps` launches on rotating frames."""
    prev = lib.cvgs_b200_set_overlap(0)
    s = wl.sets
    try:
        sp = stream.cuda_stream
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        lat = []
        for i in range(reps + 20):
            k = i % s.n
            torch.cuda.synchronize()
            e0.record(stream)
            _abi.check(lib.cvgs_b200_preproc_launch_ex(s.crops_pp[k], s.parents_pp[k], s.n_arr[k], s.n_arr[k], s.pipes_pp[k], sp))
            e1.record(stream)
            torch.cuda.synchronize()
            if i >= 20:
                lat.append(e0.elapsed_time(e1) * 1e3)
        lat.sort()
    finally:
        lib.cvgs_b200_set_overlap(prev)
    return {"what": "one isolated 50-crop launch between two events on an idle stream (includes the event overhead), plain "
                    "stream order", "latency_us_min": lat[0], "latency_us_median": lat[len(lat) // 2], "launches": len(lat)}


def c3_extra(lib, torch, _abi, marshal, util, stream, reps=100):
    """BASELINE configs[2]: 256 crops (224..896 px) of a 3840x2160 frame -> 224x224, BGR2RGB, ImageNet mean/std,
    NCHW.  154 MB of output per launch, 2 rotating (frame, tensor) sets > L2.  One launch per 256-crop batch."""
    sets = []
    for k in range(4):
        img, rects = make_c3(seed=3 + k)
        d_img = torch.from_numpy(img).cuda()
        d_out = torch.empty((256, 3, 224, 224), dtype=torch.float32, device="cuda")
        sets.append((img, rects, d_img, d_out))

    def frame_sets(n):
        return marshal.FrameSets([(s[2].data_ptr(), s[0].shape[1], 3840, 2160, s[1]) for s in sets[:n]],
                                 [s[3].data_ptr() for s in sets[:n]], (224, 224), OPS_C3)
    sp = stream.cuda_stream
    prev = lib.cvgs_b200_set_coalesce(0)  # 256 crops per launch, as the config says
    try:
        # 100 launches = 4 ms per timed call: the host's first plan (20-30 us before the first kernel starts) is 0.3 us per launch.
        # Four rotating (frame, tensor) sets: a launch rewrites the tensor written four launches earlier, so three launches out
        # of four are provably independent of everything still in flight and start without the early wait; with two sets it
        # is every second launch (reported beside it).
        us_by_sets = {}
        for n in (4, 2):
            fs = frame_sets(n)
            fs.launch_sequence(lib, 2 * n, sp)
            us_by_sets[n] = sorted(_event_us(torch, stream, lambda: fs.launch_sequence(lib, reps, sp), reps) for _ in range(3))[1]
        us = us_by_sets[4]
    finally:
        lib.cvgs_b200_set_coalesce(prev)
    img0, rects0, _d, out0 = sets[0]
    want = util.run_oracle(img0, rects0, (224, 224), OPS_C3)  # all 256 crops, once
    util.assert_bit_equal(out0.cpu().numpy(), want, "bench c3: all 256 crops vs oracle")
    b_in, b_out = algorithmic_bytes(rects0, dst=(224, 224), frame=(3840, 2160))
    gbs = (b_in + b_out) / (us * 1e-6) / 1e9
    ref_us = c3_reference_us(sets, torch, stream)
    return {"reference_fused_kernel_us_per_256_crops": ref_us,
            "reference_note": "fk::executeOperations with BATCH=128 (a template parameter capped at 255, SURVEY F7): two "
                              "launches per 256 crops, same frames, oracle/_ref/libfkref_128.so",
            "workload": "c3: 256 crops (224..896 px) from 3840x2160 -> 224x224 + BGR2RGB + mean/std + NCHW, one launch",
            "parity": "all 256 planes of the timed tensor bit-equal to the oracle",
            "timing": "median of 3 timed calls of 100 launches each over 4 rotating (frame, tensor) sets (one host thread, consecutive "
                      "launches chained by programmatic dependent launch; the early wait is dropped where the library proves the "
                      "launch independent of those still in flight)",
            "us_per_launch_two_rotating_sets": us_by_sets[2],
            "us_per_launch": us, "crops_per_s": 256 / (us * 1e-6), "algorithmic_bytes_per_launch": b_in + b_out,
            "bytes_in": b_in, "bytes_out": b_out, "achieved_gbs": gbs}


def c4_extra(torch, util, stream, depth=16, reps=20):
    """BASELINE configs[3]: CircularTensor depth 16 of 1920x1080 planes, 1080p CV_8UC3 frames (no resize), BGR2RGB +
    mean/std on the new frame; one kernel per update shifts the other 15 planes and processes the new one.  Beside it
    the reference's own fk::CircularTensor::update (oracle/_ref/libfkref_ct.so) on the same frames."""
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

        def updates():
            for i in range(reps):
                ct.update(stream, mats[i % 4], *ops)
        us = _event_us(torch, stream, updates, reps)
    finally:
        ct.close()
    plane = 12 * W * H
    alg = 3 * W * H + (depth - 1) * plane + depth * plane
    out = {"workload": f"c4: CircularTensor depth {depth}, {W}x{H} planes, 1080p frames, one kernel per update",
           "us_per_update": us, "updates_per_s": 1e6 / us, "algorithmic_bytes_per_update": alg,
           "achieved_gbs": alg / (us * 1e-6) / 1e9}
    path = os.path.join(ROOT, "oracle", "_ref", "libfkref_ct.so")
    if os.path.exists(path):
        ref = C.CDLL(path)
        ref.fkref_ct_create.restype = C.c_void_p
        ref.fkref_ct_create.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int]
        ref.fkref_ct_update.restype = C.c_int
        ref.fkref_ct_update.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_float),
                                        C.POINTER(C.c_float), C.POINTER(C.c_float), C.c_void_p]
        ref.fkref_ct_destroy.argtypes = [C.c_void_p]
        h = ref.fkref_ct_create(depth, 0, W, H)
        if h:
            f3 = lambda v: (C.c_float * 3)(*v)  # noqa: E731
            mul, sub, div = f3((1 / 255.0,) * 3), f3((0.485, 0.456, 0.406)), f3((0.229, 0.224, 0.225))

            def ref_updates(n):
                for i in range(n):
                    assert ref.fkref_ct_update(h, frames[i % 4].data_ptr(), W, H, 6144, 1, mul, sub, div, stream.cuda_stream) == 0
            ref_updates(depth + 2)
            out["reference_us_per_update"] = _event_us(torch, stream, lambda: ref_updates(reps), reps)
            out["reference_note"] = ("fk::CircularTensor<float, 3, 16, NewestFirst, Standard>::update from the reference's "
                                     "headers (circular_tensor.cuh:111-146), same frames and chain")
            torch.cuda.synchronize()
            ref.fkref_ct_destroy(h)
    return out


def c5_extra(lib, torch, dist, _abi, marshal, util, rank, world, n_total, reps=5):
    """BASELINE configs[4]: n_total crops (default 8192; w, h ~ U{224..896}) of one 4K frame -> 224x224 + BGR2RGB +
    mean/std, sharded over the ranks (contiguous ranges); every rank ends with the whole [n, 3, 224, 224] tensor.
      kernel only        each rank's slab in place (strong scaling of the compute)
      kernel + gather    then ONE in-place ncclAllGather (torch.distributed) over NVLink
      fused              the kernel stores every plane into all ranks' tensors through peer-mapped memory
                         (cvgs_b200_preproc_launch_replicated) + one 4-byte all-reduce as the completion signal
    Device time per pass, max over ranks.  Sampled planes of both tensors are compared with the oracle on every rank."""
    from cvgpuspeedup_b200 import sharding
    W = H = 224
    img, rects = make_c3(seed=5, n=n_total)  # same frame and rect list on every rank
    d_img = torch.from_numpy(img).cuda()
    lo, hi = sharding.shard_range(n_total, rank, world)
    plane_floats = 3 * W * H
    out_nccl = torch.full((n_total, 3, H, W), float("nan"), dtype=torch.float32, device="cuda")
    peer = sharding.PeerTensor((n_total, 3, H, W))
    peer.tensor.fill_(float("nan"))
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    stream = torch.cuda.current_stream()
    sp = stream.cuda_stream
    crops = marshal.crop_array(d_img.data_ptr(), img.shape[1], rects[lo:hi])
    parents = marshal.parent_array(d_img.data_ptr(), 3840, 2160, hi - lo)
    p_nccl = marshal.make_pipeline((W, H), OPS_C3, out_ptr=out_nccl[lo:hi].data_ptr())
    p_peer = marshal.make_pipeline((W, H), OPS_C3, out_ptr=peer.ptr + 4 * plane_floats * lo)
    reps_arr = (C.c_void_p * max(1, len(peer.peers)))(*[p + 4 * plane_floats * lo for p in peer.peers])
    signal = torch.zeros(1, device="cuda")
    n_mine = hi - lo

    def kernel_only():
        _abi.check(lib.cvgs_b200_preproc_launch_ex(crops, parents, n_mine, n_mine, C.byref(p_nccl), sp))

    def kernel_gather():
        kernel_only()
        sharding.gather_slabs(out_nccl, n_total)

    def fused():
        _abi.check(lib.cvgs_b200_preproc_launch_replicated(crops, parents, n_mine, n_mine, C.byref(p_peer), reps_arr,
                                                           len(peer.peers), sp))
        if dist is not None:
            dist.all_reduce(signal)

    def timed(fn):
        fn()
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(reps):
            fn()
        e1.record(stream)
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / reps], device="cuda", dtype=torch.float64)
        if dist is not None:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    ms_k = timed(kernel_only)
    ms_g = timed(kernel_gather) if world > 1 else ms_k
    ms_f = timed(fused) if world > 1 else None
    # parity: planes of every rank's slab, on every rank, both tensors
    idx = sorted({0, n_total - 1, lo, hi - 1, n_total // 2, n_total // 3, (lo + hi) // 2,
                  *[sharding.shard_range(n_total, r, world)[0] for r in range(world)]})
    want = util.run_oracle(img, [rects[i] for i in idx], (W, H), OPS_C3)
    if world > 1:
        util.assert_bit_equal(out_nccl[idx].cpu().numpy(), want, f"c5 rank {rank}: all-gathered tensor vs oracle")
        util.assert_bit_equal(peer.tensor[idx].cpu().numpy(), want, f"c5 rank {rank}: peer-stored tensor vs oracle")
    else:
        mine = [i for i in idx if lo <= i < hi]
        util.assert_bit_equal(out_nccl[mine].cpu().numpy(), util.run_oracle(img, [rects[i] for i in mine], (W, H), OPS_C3),
                              "c5: tensor vs oracle")
    peer.close()
    total_bytes = 4 * plane_floats * n_total
    recv = total_bytes * (world - 1) / world  # bytes every GPU receives over NVLink
    out = {"workload": f"c5: {n_total} crops (224..896 px) of one 3840x2160 frame -> 224x224 + BGR2RGB + mean/std, "
                       f"contiguous ranges of {n_total // world} crops per GPU, whole [n,3,224,224] tensor "
                       f"({total_bytes / 1e9:.2f} GB) on every GPU",
           "n_gpus": world, "kernel_only_ms": ms_k, "kernel_only_crops_per_s": n_total / (ms_k * 1e-3),
           "parity": f"planes {idx} bit-equal to the oracle on every rank (both tensors)"}
    if world > 1:
        out.update({"kernel_plus_allgather_ms": ms_g, "kernel_plus_allgather_crops_per_s": n_total / (ms_g * 1e-3),
                    "allgather_nvlink_gbs_per_gpu_in": recv / ((ms_g - ms_k) * 1e-3) / 1e9 if ms_g > ms_k else None,
                    "fused_peer_store_ms": ms_f, "fused_peer_store_crops_per_s": n_total / (ms_f * 1e-3),
                    "fused_nvlink_gbs_per_gpu_in": recv / (ms_f * 1e-3) / 1e9,
                    "nvlink_bound_ms_at_770_gbs": recv / 770e9 * 1e3,
                    "note": "every GPU must receive (N-1)/N of the tensor: the NVLink inbound rate bounds both variants; "
                            "770 GB/s per direction is the pool's measured peer-copy rate (B200_PROFILING.md)"})
    return out


def _fkref_lib():
    path = os.path.join(ROOT, "oracle", "_ref", "libfkref_50.so")
    return C.CDLL(path) if os.path.exists(path) else None


def gpu_baselines(frames, d_imgs, torch, min_seconds=0.5):
    """Reference fused kernel (its own headers, BATCH=50 instantiation) on the same device frames: a native frame loop
    from one host thread in stream order (how the reference's API is used) and from three threads / streams (the loop
    round 1's headline used), so that kernel and launch strategy can be told apart."""
    out = {}
    lib = _fkref_lib()
    if lib is not None:
        P, I = C.POINTER, C.c_int
        f3 = lambda v: (C.c_float * 3)(*v)  # noqa: E731
        n = CROPS_PER_FRAME
        argsets = []
        outs = [torch.empty((n, 3, DST[1], DST[0]), dtype=torch.float32, device="cuda") for _ in frames]
        for (img, rects), d in zip(frames, d_imgs):
            base = d.data_ptr()
            ptrs = (C.c_void_p * n)(*[base + y * PITCH + 3 * x for (x, y, w, h) in rects])
            argsets.append((ptrs, (I * n)(*[r[2] for r in rects]), (I * n)(*[r[3] for r in rects]), (I * n)(*[PITCH] * n)))
        bg, mul, sub, div = f3((0, 0, 0)), f3(MUL), f3(SUB), f3(DIV)
        s = torch.cuda.current_stream()
        nf = len(frames)
        a_ptrs = (P(C.c_void_p) * nf)(*[C.cast(a[0], P(C.c_void_p)) for a in argsets])
        a_ws = (P(I) * nf)(*[C.cast(a[1], P(I)) for a in argsets])
        a_hs = (P(I) * nf)(*[C.cast(a[2], P(I)) for a in argsets])
        a_ps = (P(I) * nf)(*[C.cast(a[3], P(I)) for a in argsets])
        a_outs = (C.c_void_p * nf)(*[o.data_ptr() for o in outs])
        what = ("fk::executeOperations(BatchRead<50>(Resize<INTER_LINEAR>), ColorConversion, Mul, Sub, Div, TensorSplit) "
                "from /root/reference/fkl/include compiled for sm_100a (oracle/_ref/libfkref_50.so), same device frames, ")
        for key, fname, threads in (("reference_fused_kernel_gpu", "fkref_preproc_sequence_50", 1),
                                    ("reference_fused_kernel_gpu_3_threads", "fkref_preproc_sequence_mt_50", 3)):
            seq = getattr(lib, fname, None)
            if seq is None:
                continue
            seq.restype = I
            base_types = [P(P(C.c_void_p)), P(P(I)), P(P(I)), P(P(I)), I, I, I, I, P(C.c_float), I, P(C.c_float),
                          P(C.c_float), P(C.c_float), P(C.c_void_p), I, I, C.c_void_p]
            seq.argtypes = base_types + ([I] if threads > 1 else [])

            def passes(k, seq=seq, threads=threads):
                a = [a_ptrs, a_ws, a_hs, a_ps, n, DST[0], DST[1], 1, bg, 1, mul, sub, div, a_outs, nf, nf * k, s.cuda_stream]
                assert seq(*(a + ([threads] if threads > 1 else []))) == 0
            passes(20)
            reps = 50
            us = _event_us(torch, s, lambda: passes(reps), reps * nf)
            out[key] = {"value": n / (us * 1e-6), "unit": "crops/s", "us_per_launch": us, "host_threads": threads,
                        "what": what + (f"native frame loop, {threads} host threads with one stream each" if threads > 1
                                        else "native frame loop from one host thread, stream order")}
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
    for (img, rects, d_img, d_out) in sets:
        ref_out = torch.empty_like(d_out)
        pitch = img.shape[1]
        for half in range(2):
            rr = rects[128 * half:128 * (half + 1)]
            ptrs = (C.c_void_p * 128)(*[d_img.data_ptr() + y * pitch + 3 * x for (x, y, _, _) in rr])
            calls.append((ptrs, (C.c_int * 128)(*[r[2] for r in rr]), (C.c_int * 128)(*[r[3] for r in rr]),
                          (C.c_int * 128)(*[pitch] * 128), ref_out[128 * half:].data_ptr(), ref_out))
    mul, sub, div = (1 / 255.0,) * 3, (0.485, 0.456, 0.406), (0.229, 0.224, 0.225)

    def one_pass():
        for ptrs, ws, hs, ps, optr, _keep in calls:
            rc = fn(ptrs, ws, hs, ps, 128, 224, 224, 1, f3((0, 0, 0)), 1, f3(mul), f3(sub), f3(div), optr, stream.cuda_stream)
            assert rc == 0
    one_pass()

    def passes():
        for _ in range(reps):
            one_pass()
    return _event_us(torch, stream, passes, reps * len(sets))


def h2d_bytes(frames) -> int:
    """Bytes cvgs_b200_preproc_host would upload per step with plain row copies: rows [min y, max y+h) of each frame."""
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
    ap.add_argument("--no-extras", action="store_true", help="skip the c2 variants and c3 / c4 / c5")
    ap.add_argument("--c5-crops", type=int, default=8192)
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
