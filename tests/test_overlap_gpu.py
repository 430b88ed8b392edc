"""cvgs_b200_set_overlap(1 / 2): consecutive launches may overlap when the library proves them independent.  Stream
semantics must be unchanged: hazards between launches (same output, output used as a source) are ordered, and
whatever follows on the stream sees every earlier launch complete."""
import ctypes as C

import numpy as np
import pytest
import torch

from cvgpuspeedup_b200 import _abi
from tests import util

pytestmark = pytest.mark.gpu


@pytest.fixture()
def overlap_on():
    lib = _abi.load()
    prev = lib.cvgs_b200_set_overlap(2)  # mode 2: individual launches may overlap too (the tests enqueue nothing foreign in between)
    yield lib
    lib.cvgs_b200_set_overlap(prev)


def _launch(lib, w, d_img, out, stream, parents=True, rects=None):
    rects = w.rects if rects is None else rects
    crops = util.host_crops(w.image, rects, base_ptr=d_img.data_ptr())
    p = util.make_pipeline(w.dsize, w.ops, out_ptr=out.data_ptr())
    par = util.host_parents(w.image, w.width, w.height, len(rects), base_ptr=d_img.data_ptr()) if parents else None
    _abi.check(lib.cvgs_b200_preproc_launch_ex(crops, par, len(rects), len(rects), C.byref(p), stream.cuda_stream))
    return crops, p, par  # keep alive until the caller synchronises


def test_many_independent_launches(overlap_on):
    lib = overlap_on
    st = torch.cuda.Stream()
    ws = [util.workload_c2(seed=300 + k, n=20, frame=(640, 480), pitch=1920) for k in range(6)]
    d_imgs = [torch.from_numpy(w.image).cuda() for w in ws]
    outs = [torch.full((20, 3, 128, 64), float("nan"), device="cuda") for _ in range(40)]
    torch.cuda.synchronize()  # the fills ran on the default stream, the launches go to a non-blocking one
    keep = [_launch(lib, ws[i % 6], d_imgs[i % 6], outs[i], st) for i in range(40)]  # > the 8-launch window
    st.synchronize()
    want = [util.run_oracle(w.image, w.rects, w.dsize, w.ops) for w in ws]
    for i, o in enumerate(outs):
        util.assert_bit_equal(o.cpu().numpy(), want[i % 6], f"launch {i}")
    del keep


def test_same_output_is_written_in_launch_order(overlap_on):
    lib = overlap_on
    st = torch.cuda.Stream()
    ws = [util.workload_c2(seed=310 + k, n=50, pitch=6144) for k in range(2)]
    d_imgs = [torch.from_numpy(w.image).cuda() for w in ws]
    out = torch.full((50, 3, 128, 64), float("nan"), device="cuda")
    torch.cuda.synchronize()  # the fills ran on the default stream, the launches go to a non-blocking one
    keep = []
    for rep in range(25):
        for k in range(2):
            keep.append(_launch(lib, ws[k], d_imgs[k], out, st))
    st.synchronize()
    util.assert_bit_equal(out.cpu().numpy(), util.run_oracle(ws[1].image, ws[1].rects, ws[1].dsize, ws[1].ops), "last writer wins")


def test_output_of_one_launch_as_source_of_the_next(overlap_on):
    """Launch B reads (as CV_8UC3 bytes) the float tensor launch A writes: a read-after-write hazard through memory."""
    lib = overlap_on
    st = torch.cuda.Stream()
    wa = util.workload_c3(seed=320, n=64, frame=(1920, 1080), dsize=(224, 224), lo=200, hi=800)
    d_img = torch.from_numpy(wa.image).cuda()
    t = torch.zeros((64, 3, 224, 224), dtype=torch.float32, device="cuda")      # A's output = B's "image"
    torch.cuda.synchronize()  # the fills ran on the default stream, the launches go to a non-blocking one
    pitch = 3 * 224 * 4                                                          # one float row triple = 896 pixels
    rows = t.numel() * 4 // pitch
    rects = [(5 * i, 3 * i, 200 + i, 100 + 2 * i) for i in range(30)]
    out_b = torch.full((30, 3, 128, 64), float("nan"), device="cuda")
    torch.cuda.synchronize()  # the fills ran on the default stream, the launches go to a non-blocking one
    keep = []
    for rep in range(3):
        keep.append(_launch(lib, wa, d_img, t, st))
        crops = (_abi.Crop * 30)()
        for i, (x, y, w, h) in enumerate(rects):
            crops[i].data, crops[i].width, crops[i].height, crops[i].pitch = t.data_ptr() + y * pitch + 3 * x, w, h, pitch
        p = util.make_pipeline((64, 128), util.OPS_C2, out_ptr=out_b.data_ptr())
        _abi.check(lib.cvgs_b200_preproc_launch_ex(crops, None, 30, 30, C.byref(p), st.cuda_stream))
        keep.append((crops, p))
    st.synchronize()
    host_t = t.cpu().numpy().view(np.uint8).reshape(rows, pitch)
    util.assert_bit_equal(t.cpu().numpy(), util.run_oracle(wa.image, wa.rects, wa.dsize, wa.ops), "launch A")
    util.assert_bit_equal(out_b.cpu().numpy(), util.run_oracle(host_t, rects, (64, 128), util.OPS_C2), "launch B sees A's output")


def test_later_stream_work_sees_every_launch_complete(overlap_on):
    """A long launch followed by a short independent one, then a device-to-host copy of the FIRST launch's output on
    the same stream: the copy must see it complete (each kernel waits for its predecessor before it exits)."""
    lib = overlap_on
    st = torch.cuda.Stream()
    big = util.workload_c3(seed=330, n=256)
    small = util.workload_c2(seed=331, n=2, frame=(320, 240), pitch=960)
    d_big, d_small = torch.from_numpy(big.image).cuda(), torch.from_numpy(small.image).cuda()
    out_big = torch.full((256, 3, 224, 224), float("nan"), device="cuda")
    out_small = torch.full((2, 3, 128, 64), float("nan"), device="cuda")
    torch.cuda.synchronize()  # the fills ran on the default stream, the launches go to a non-blocking one
    host = torch.empty(out_big.shape, dtype=torch.float32).pin_memory()
    want = util.run_oracle(big.image, big.rects[:8], big.dsize, big.ops)
    for rep in range(5):
        out_big.fill_(float("nan"))
        torch.cuda.synchronize()
        k1 = _launch(lib, big, d_big, out_big, st)
        k2 = _launch(lib, small, d_small, out_small, st)
        with torch.cuda.stream(st):
            host.copy_(out_big, non_blocking=True)
        st.synchronize()
        assert not np.isnan(host.numpy()).any(), "copy overtook the first launch"
        util.assert_bit_equal(host.numpy()[:8], want, "first launch output")
        del k1, k2


@pytest.mark.parametrize("overlap", [0, 1, 2])
def test_launches_are_graph_capturable(overlap):
    """Batches whose descriptors ride in the kernel parameters involve no copy and no allocation on the hot call, so a
    frame loop can be captured into a CUDA graph and replayed (graph replay figure for config 2)."""
    lib = _abi.load()
    prev = lib.cvgs_b200_set_overlap(overlap)
    try:
        ws = [util.workload_c2(seed=340 + k, n=50, pitch=6144) for k in range(3)]
        d_imgs = [torch.from_numpy(w.image).cuda() for w in ws]
        outs = [torch.full((50, 3, 128, 64), float("nan"), device="cuda") for _ in ws]
        torch.cuda.synchronize()  # the fills ran on the default stream, the launches go to a non-blocking one
        st = torch.cuda.Stream()
        keep = [_launch(lib, w, d, o, st) for w, d, o in zip(ws, d_imgs, outs)]  # warm-up: attributes, proofs, map cache
        st.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=st):
            keep += [_launch(lib, w, d, o, torch.cuda.current_stream()) for w, d, o in zip(ws, d_imgs, outs)]
        for rep in range(3):
            for o in outs:
                o.fill_(float("nan"))
            g.replay()
            torch.cuda.synchronize()
            for w, o in zip(ws, outs):
                util.assert_bit_equal(o.cpu().numpy(), util.run_oracle(w.image, w.rects, w.dsize, w.ops), f"replay {rep}")
    finally:
        lib.cvgs_b200_set_overlap(prev)


def test_concurrent_host_threads():
    """Four host threads launching on their own streams at the same time (ctypes releases the GIL): per-thread
    contexts (descriptor ring, tensor-map cache, memos) and the shared bookkeeping must not interfere."""
    import threading
    lib = _abi.load()
    prev = lib.cvgs_b200_set_overlap(2)
    errors = []

    def worker(k):
        try:
            torch.cuda.set_device(0)
            st = torch.cuda.Stream()
            small = util.workload_c2(seed=400 + k, n=30, frame=(640, 480), pitch=1920)
            big = util.workload_c2(seed=500 + k, n=100, frame=(640, 480), pitch=1920)   # > 64 crops: 256-crop table
            ring = util.workload_c2(seed=600 + k, n=300, frame=(640, 480), pitch=1920)  # > 256 crops: descriptor ring
            for w in (small, big, ring):
                w.dsize = (32, 48)
            ds = [torch.from_numpy(w.image).cuda() for w in (small, big, ring)]
            wants = [util.run_oracle(w.image, w.rects, w.dsize, w.ops) for w in (small, big, ring)]
            for rep in range(15):
                outs, keep = [], []
                for w, d in zip((small, big, ring), ds):
                    with torch.cuda.stream(st):  # the fill is ordered in front of the launch (st does not sync with the default stream)
                        o = torch.full((len(w.rects), 3, 48, 32), float("nan"), device="cuda")
                    keep.append(_launch(lib, w, d, o, st))
                    outs.append(o)
                st.synchronize()
                for o, want in zip(outs, wants):
                    with torch.cuda.stream(st):
                        got = o.cpu().numpy()
                    util.assert_bit_equal(got, want, f"thread {k} rep {rep}")
        except Exception as e:  # noqa: BLE001
            errors.append(f"thread {k}: {e}")

    threads = [threading.Thread(target=worker, args=(k,)) for k in range(4)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    lib.cvgs_b200_set_overlap(prev)
    assert not errors, errors


def _sequence(lib, ws, d_imgs, outs, steps, stream):
    n = len(ws)
    keep = []
    for w, d, o in zip(ws, d_imgs, outs):
        crops = util.host_crops(w.image, w.rects, base_ptr=d.data_ptr())
        par = util.host_parents(w.image, w.width, w.height, len(w.rects), base_ptr=d.data_ptr())
        p = util.make_pipeline(w.dsize, w.ops, out_ptr=o.data_ptr())
        keep.append((crops, par, p))
    crops_pp = (C.POINTER(_abi.Crop) * n)(*[C.cast(k[0], C.POINTER(_abi.Crop)) for k in keep])
    par_pp = (C.POINTER(_abi.Parent) * n)(*[C.cast(k[1], C.POINTER(_abi.Parent)) for k in keep])
    pipes_pp = (C.POINTER(_abi.Pipeline) * n)(*[C.pointer(k[2]) for k in keep])
    n_arr = (C.c_int32 * n)(*[len(w.rects) for w in ws])
    before = lib.cvgs_b200_launch_count()
    _abi.check(lib.cvgs_b200_preproc_launch_sequence_ex(crops_pp, par_pp, n_arr, n_arr, pipes_pp, n, steps,
                                                        stream.cuda_stream))
    return lib.cvgs_b200_launch_count() - before, keep


@pytest.fixture()
def per_step_launches(overlap_on):
    """cvgs_b200_set_coalesce(0): the frame loop launches once per step (helper threads) instead of sharing launches."""
    prev = overlap_on.cvgs_b200_set_coalesce(0)
    yield overlap_on
    overlap_on.cvgs_b200_set_coalesce(prev)


def test_frame_loop_shares_launches(overlap_on):
    """Independent argument sets with the same pipeline: consecutive steps ride in shared launches (up to 512 crops of
    up to 32 sets), every crop landing in its own set's tensor; results, ordering on the caller's stream and the
    reduced launch count hold.  Sets of different sizes, more steps than sets (tensors are rewritten: write-after-
    write hazards between launches), and a run long enough to wrap the in-flight window."""
    lib = overlap_on
    st = torch.cuda.Stream()
    sizes = [50, 50, 7, 50, 120, 50, 1, 50, 33, 50, 50]
    ws = [util.workload_c2(seed=740 + k, n=n, pitch=6144) for k, n in enumerate(sizes)]
    d_imgs = [torch.from_numpy(w.image).cuda() for w in ws]
    outs = [torch.full((len(w.rects), 3, 128, 64), float("nan"), device="cuda") for w in ws]
    torch.cuda.synchronize()  # the fills ran on the default stream, the launches go to a non-blocking one
    want = [util.run_oracle(w.image, w.rects, w.dsize, w.ops) for w in ws]
    steps = len(ws) * 9 + 5
    with torch.cuda.stream(st):
        launches, keep = _sequence(lib, ws, d_imgs, outs, steps, st)
        copies = [o.clone() for o in outs]
    st.synchronize()
    assert 0 < launches <= steps // 5, launches  # 511 crops per pass over the sets: about one launch per pass
    for k, (o, c) in enumerate(zip(outs, copies)):
        util.assert_bit_equal(o.cpu().numpy(), want[k], f"set {k}")
        util.assert_bit_equal(c.cpu().numpy(), want[k], f"set {k}: copy queued after the loop")
    # a second sequence on the same stream reuses the cached tensor maps
    for o in outs:
        o.fill_(float("nan"))
    torch.cuda.synchronize()  # the fills ran on the default stream, the loop runs on st
    launches2, keep2 = _sequence(lib, ws, d_imgs, outs, len(ws), st)
    st.synchronize()
    assert launches2 <= 2
    for k, o in enumerate(outs):
        util.assert_bit_equal(o.cpu().numpy(), want[k], f"second sequence, set {k}")
    del keep, keep2


def test_frame_loop_shared_launches_overlap_and_rewrite(overlap_on):
    """More sets than one launch takes (24 x 50 crops: 10 sets per launch) and several passes over them: consecutive
    launches overlap (late griddepcontrol.wait) until one rewrites a tensor a launch possibly in flight wrote.  Between
    passes the sources of half the sets change (device-side copy on the same stream), so a launch that ran ahead of the
    stream order or a stale tensor would show."""
    lib = overlap_on
    st = torch.cuda.Stream()
    n_sets = 24
    ws_a = [util.workload_c2(seed=800 + k, n=50, frame=(960, 540), pitch=2880) for k in range(n_sets)]
    ws_b = [util.workload_c2(seed=900 + k, n=50, frame=(960, 540), pitch=2880) for k in range(n_sets)]
    for a, b in zip(ws_a, ws_b):
        b.rects = a.rects  # same crops, other pixels
    d_imgs = [torch.from_numpy(w.image).cuda() for w in ws_a]
    d_alt = [torch.from_numpy(w.image).cuda() for w in ws_b]
    outs = [torch.full((50, 3, 128, 64), float("nan"), device="cuda") for _ in ws_a]
    torch.cuda.synchronize()  # the fills ran on the default stream, the launches go to a non-blocking one
    with torch.cuda.stream(st):
        launches, keep = _sequence(lib, ws_a, d_imgs, outs, 3 * n_sets + 7, st)
        for k in range(0, n_sets, 2):
            d_imgs[k].copy_(d_alt[k], non_blocking=True)   # stream-ordered behind the first loop
        launches2, keep2 = _sequence(lib, ws_a, d_imgs, outs, 2 * n_sets, st)
    st.synchronize()
    assert launches <= 10 and launches2 <= 6, (launches, launches2)
    for k in range(n_sets):
        w = ws_b[k] if k % 2 == 0 else ws_a[k]
        util.assert_bit_equal(outs[k].cpu().numpy(), util.run_oracle(w.image, w.rects, w.dsize, w.ops), f"set {k}")
    del keep, keep2


def test_frame_loop_shared_launch_c3_shape(overlap_on):
    """224x224 planes (two column bands of different width) from three 4K frames in shared launches."""
    lib = overlap_on
    st = torch.cuda.Stream()
    ws = [util.workload_c3(seed=760 + k, n=40) for k in range(3)]
    d_imgs = [torch.from_numpy(w.image).cuda() for w in ws]
    outs = [torch.full((40, 3, 224, 224), float("nan"), device="cuda") for _ in ws]
    torch.cuda.synchronize()  # the fills ran on the default stream, the launches go to a non-blocking one
    launches, keep = _sequence(lib, ws, d_imgs, outs, 6, st)
    st.synchronize()
    assert launches == 2
    for k, (w, o) in enumerate(zip(ws, outs)):
        util.assert_bit_equal(o.cpu().numpy(), util.run_oracle(w.image, w.rects, w.dsize, w.ops), f"set {k}")
    del keep


def test_frame_loop_declines_to_share_what_the_tma_kernel_cannot_take(overlap_on):
    """A set whose frame pitch is not a multiple of 16 bytes cannot be staged by TMA: the group it falls into is
    launched set by set (direct-gather kernel for that set), results unchanged."""
    lib = overlap_on
    st = torch.cuda.Stream()
    ws = [util.workload_c2(seed=770 + k, n=20, frame=(640, 480), pitch=1920 if k != 2 else 1923) for k in range(5)]
    d_imgs = [torch.from_numpy(w.image).cuda() for w in ws]
    outs = [torch.full((20, 3, 128, 64), float("nan"), device="cuda") for _ in ws]
    torch.cuda.synchronize()  # the fills ran on the default stream, the launches go to a non-blocking one
    launches, keep = _sequence(lib, ws, d_imgs, outs, 10, st)
    st.synchronize()
    assert launches == 10
    for k, (w, o) in enumerate(zip(ws, outs)):
        util.assert_bit_equal(o.cpu().numpy(), util.run_oracle(w.image, w.rects, w.dsize, w.ops), f"set {k}")
    del keep


def test_frame_loop_over_several_host_threads(per_step_launches):
    """steps >= 16 with independent argument sets: the loop is split over helper threads and streams
    (include/cvgs_b200.h); every frame's result, the launch count and the ordering on the caller's stream hold."""
    lib = per_step_launches
    st = torch.cuda.Stream()
    ws = [util.workload_c2(seed=700 + k, n=50, pitch=6144) for k in range(7)]
    d_imgs = [torch.from_numpy(w.image).cuda() for w in ws]
    outs = [torch.full((50, 3, 128, 64), float("nan"), device="cuda") for _ in ws]
    torch.cuda.synchronize()  # the fills ran on the default stream, the launches go to a non-blocking one
    want = [util.run_oracle(w.image, w.rects, w.dsize, w.ops) for w in ws]
    with torch.cuda.stream(st):
        launches, keep = _sequence(lib, ws, d_imgs, outs, 7 * 30, st)
        # work queued behind the loop on the caller's stream must see every frame complete, whichever stream ran it
        copies = [o.clone() for o in outs]
    st.synchronize()
    assert launches == 7 * 30
    for k, (o, c) in enumerate(zip(outs, copies)):
        util.assert_bit_equal(o.cpu().numpy(), want[k], f"set {k}")
        util.assert_bit_equal(c.cpu().numpy(), want[k], f"set {k}: copy queued after the loop")
    del keep


def test_frame_loop_with_dependent_sets_stays_in_order(overlap_on):
    """Two argument sets that write the same tensor are not independent: the loop must run in plain order, so the
    last call (set 1) wins; a set whose source is another set's output likewise."""
    lib = overlap_on
    st = torch.cuda.Stream()
    ws = [util.workload_c2(seed=720 + k, n=50, pitch=6144) for k in range(2)]
    d_imgs = [torch.from_numpy(w.image).cuda() for w in ws]
    out = torch.full((50, 3, 128, 64), float("nan"), device="cuda")
    torch.cuda.synchronize()  # the fills ran on the default stream, the launches go to a non-blocking one
    launches, keep = _sequence(lib, ws, d_imgs, [out, out], 2 * 80, st)
    st.synchronize()
    assert launches == 160
    util.assert_bit_equal(out.cpu().numpy(), util.run_oracle(ws[1].image, ws[1].rects, ws[1].dsize, ws[1].ops), "last writer wins")
    del keep


@pytest.mark.parametrize("coalesce", [0, 1])
def test_frame_loop_error_is_reported(overlap_on, coalesce):
    lib = overlap_on
    lib.cvgs_b200_set_coalesce(coalesce)
    st = torch.cuda.Stream()
    ws = [util.workload_c2(seed=730 + k, n=10, frame=(640, 480), pitch=1920) for k in range(4)]
    d_imgs = [torch.from_numpy(w.image).cuda() for w in ws]
    outs = [torch.zeros((10, 3, 128, 64), device="cuda") for _ in ws]
    torch.cuda.synchronize()  # the fills ran on the default stream, the launches go to a non-blocking one
    n = 4
    keep = []
    for i, (w, d, o) in enumerate(zip(ws, d_imgs, outs)):
        crops = util.host_crops(w.image, w.rects, base_ptr=d.data_ptr())
        if i == 3:
            crops[2].pitch = 1  # smaller than a row: rejected when that set is launched (by a helper thread)
        par = util.host_parents(w.image, w.width, w.height, len(w.rects), base_ptr=d.data_ptr())
        keep.append((crops, par, util.make_pipeline(w.dsize, w.ops, out_ptr=o.data_ptr())))
    crops_pp = (C.POINTER(_abi.Crop) * n)(*[C.cast(k[0], C.POINTER(_abi.Crop)) for k in keep])
    par_pp = (C.POINTER(_abi.Parent) * n)(*[C.cast(k[1], C.POINTER(_abi.Parent)) for k in keep])
    pipes_pp = (C.POINTER(_abi.Pipeline) * n)(*[C.pointer(k[2]) for k in keep])
    n_arr = (C.c_int32 * n)(*[10] * n)
    rc = lib.cvgs_b200_preproc_launch_sequence_ex(crops_pp, par_pp, n_arr, n_arr, pipes_pp, n, 400, st.cuda_stream)
    st.synchronize()
    lib.cvgs_b200_set_coalesce(1)
    assert rc != 0 and b"pitch" in lib.cvgs_b200_last_error()
