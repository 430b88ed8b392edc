// preproc.cu -- the fused crop -> bilinear resize -> op chain -> planar split path for sm_100a.
//
// Replaces the one instantiation of fk::launchTransformDPP_Kernel that
// cvGS::executeOperations(stream, resize<CV_8UC3,...>(crops...), ops..., split<CV_32FC3>(...)) produces
// (reference fkl/include/fused_kernel/core/execution_model/data_parallel_patterns.cuh:157-197,256-260;
// launch logic executors.cuh:109-158).  Batch size, aspect mode and the op chain are runtime data
// here (they are template parameters in the reference, capped at 255 planes -- SURVEY.md F7).
//
// Two kernels:
//   preproc_tma_kernel     (preproc_tma.cuh) the product path: persistent, every warp stages the source rows of its
//                          own work items into shared memory with the TMA engine (cp.async.bulk.tensor, mbarrier
//                          per slot) and computes row pairs in packed FP32.
//   preproc_direct_kernel  gathers the bilinear taps straight from global memory; takes anything the TMA kernel
//                          declines: pitches that are not multiples of 16 bytes, extreme down-scales, 16-bit and
//                          4-channel sources, YUV frames, conversions that change the channel count.
// This file: the direct kernel, the host launch path (descriptor tables in kernel parameters or through a pinned
// ring, per-image tensor-map cache, overlap bookkeeping) and the C-ABI entry points of the batch pipeline.
#include <cuda_runtime.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <cstdlib>
#include <chrono>
#include <condition_variable>
#include <functional>
#include <cstdio>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "cvgs_device.cuh"
#include "cvgs_runtime.hpp"
#include "preproc_host.hpp"
#include "preproc_direct.cuh"
#include "preproc_tma.cuh"
#include "preproc_yuv_tma.cuh"
#include "preproc_warp.cuh"

namespace cvgs {

// ------------------------------------------------------------------------------------------------
// errors
// ------------------------------------------------------------------------------------------------
static thread_local std::string t_last_error;
// host-side cost breakdown of the launch path (diagnostics, cvgs_b200_debug_host_profile)
struct HostProfile {
    double fill = 0, plan = 0, encode = 0, launch = 0;
    long long calls = 0;
};
static thread_local HostProfile t_prof;
static inline double now_us() {
    return std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now().time_since_epoch()).count();
}
static thread_local int64_t t_launch_count = 0;
static std::atomic<int> g_variant{0};
static std::atomic<int> g_host_tiles{[] {
    const char* e = std::getenv("CVGS_B200_HOST_TILES");
    return e && e[0] == '1' ? 1 : 0;
}()};
static std::atomic<int> g_coalesce{[] {
    const char* e = std::getenv("CVGS_B200_SEQ_COALESCE");
    return e && e[0] == '0' ? 0 : 1;
}()};
// 0 off; 1 = inside the library's own frame loops, where it controls the whole stream segment; 2 = also between
// individual launches (the caller asserts that nothing it enqueues between two launches of the library produces a source)
static std::atomic<int> g_overlap{[] {
    const char* e = std::getenv("CVGS_B200_OVERLAP");
    return e && e[0] == '1' ? 1 : (e && e[0] == '2' ? 2 : 0);
}()};
static thread_local bool t_in_frame_loop = false;  // set by the multi-threaded frame loop around the launches it issues

// ------------------------------------------------------------------------------------------------
// Overlap of consecutive launches (cvgs_b200_set_overlap).  The TMA kernel is launched with programmatic stream
// serialisation and normally orders itself after the preceding kernel with griddepcontrol.wait before its first
// global access.  With overlap enabled the library keeps, per stream, the memory ranges of its recent launches
// and drops that early wait when the new launch neither reads nor writes what they write and does not write what
// they read; every kernel still waits for its predecessor before it exits, so completion order (what any later
// stream operation observes) is unchanged.  At most kWindow launches can be in flight this way.
// ------------------------------------------------------------------------------------------------
struct MemRange {
    uintptr_t lo = 0, hi = 0;
    bool overlaps(const MemRange& o) const { return lo < o.hi && o.lo < hi; }
};
struct StreamTrack {
    cudaStream_t stream = nullptr;
    bool used = false;
    static constexpr int kWindow = 8;
    MemRange out[kWindow], src[kWindow];
    int n = 0;
    bool unknown_pred = false;  // something this bookkeeping did not see went to the stream: the next launch waits up front
};
static std::mutex g_track_mu;
static StreamTrack g_tracks[16];

// set by the multi-threaded frame loop: several launches are resident at once, so fewer, longer-lived CTAs per launch
// measured best there (8 items per warp: 2.13 us per 50-crop frame; 4: 2.32; 16: 2.20)
static thread_local int t_items_hint = 0;
// Items per warp small launches aim for: 1 in plain stream order (latency), more when launches overlap (throughput).
static int items_per_warp() {
    static const int forced = [] {
        const char* e = std::getenv("CVGS_TMA_ITEMS_PER_WARP");
        return e ? std::max(1, std::atoi(e)) : 0;
    }();
    if (forced) return forced;
    if (t_items_hint) return t_items_hint;
    return g_overlap.load(std::memory_order_relaxed) == 2 ? 4 : 1;
}

// true: this launch must wait for the preceding kernel up front.
static bool overlap_needs_wait(cudaStream_t stream, const MemRange& out, const MemRange& src) {
    const int mode = g_overlap.load(std::memory_order_relaxed);
    // Between individual launches the library cannot see what else the caller enqueued (a foreign kernel that produces
    // the next source would be ordered only by the implicit trigger at its exit, which guarantees nothing about the
    // visibility of its writes without a griddepcontrol.wait): the early wait is dropped there only on the caller's word.
    if (mode == 0 || (mode == 1 && !t_in_frame_loop)) return true;
    std::lock_guard<std::mutex> lock(g_track_mu);
    StreamTrack* t = nullptr;
    for (auto& k : g_tracks)
        if (k.used && k.stream == stream) { t = &k; break; }
    if (!t) {
        for (auto& k : g_tracks)
            if (!k.used) { t = &k; break; }
        if (!t) return true;  // more concurrent streams than slots: plain stream order
        t->used = true;
        t->stream = stream;
        t->n = 0;
        // first launch seen on this stream: whatever precedes it is unknown
        t->out[0] = out; t->src[0] = src; t->n = 1;
        return true;
    }
    bool hazard = t->n >= StreamTrack::kWindow || t->unknown_pred;
    t->unknown_pred = false;
    for (int i = 0; i < t->n && !hazard; ++i)
        hazard = out.overlaps(t->out[i]) || src.overlaps(t->out[i]) || out.overlaps(t->src[i]);
    if (hazard) t->n = 0;  // the wait orders this launch after everything before it
    t->out[t->n] = out;
    t->src[t->n] = src;
    ++t->n;
    return hazard;
}
static void overlap_forget(cudaStream_t stream) {  // a launch of this library that is not PDL-aware went to `stream`
    if (!g_overlap.load(std::memory_order_relaxed)) return;
    std::lock_guard<std::mutex> lock(g_track_mu);
    for (auto& k : g_tracks)
        if (k.used && k.stream == stream) {
            k.n = 0;
            k.unknown_pred = true;
        }
}
static MemRange crops_range(const DevCrop* c, int n, int pb = 3) {
    MemRange r;
    r.lo = ~static_cast<uintptr_t>(0);
    for (int i = 0; i < n; ++i) {
        const uintptr_t lo = reinterpret_cast<uintptr_t>(c[i].data);
        const uintptr_t hi = lo + static_cast<uintptr_t>(c[i].h - 1) * static_cast<uintptr_t>(c[i].pitch) + static_cast<uintptr_t>(pb) * c[i].w;
        r.lo = std::min(r.lo, lo);
        r.hi = std::max(r.hi, hi);
    }
    if (n <= 0) r.lo = r.hi = 0;
    return r;
}
static MemRange out_range(const PreprocParams& P) {
    // every layout stays inside [base, base + extent): NCHW/NHWC planes z*z_stride + 3*W*H, CNHW c*c_stride + ...
    const long long plane = static_cast<long long>(P.W) * P.H;
    MemRange r;
    r.lo = reinterpret_cast<uintptr_t>(P.out.base);
    if (P.out.u8) {  // strides in bytes
        r.hi = r.lo + static_cast<uintptr_t>((P.n_planes - 1) * P.out.z_stride + P.out.row_pitch * P.H);
        return r;
    }
    long long extent;
    const int nco = P.prog.nc_out > 0 ? P.prog.nc_out : 3;
    if (P.out.px_stride != 1) extent = (P.n_planes - 1) * P.out.z_stride + P.out.row_stride * P.H;
    else extent = (nco - 1) * P.out.c_stride + (P.n_planes - 1) * P.out.z_stride + plane;
    r.hi = r.lo + static_cast<uintptr_t>(extent) * sizeof(float);
    return r;
}

void set_error(const std::string& msg) { t_last_error = msg; }
int fail(int code, const std::string& msg) {
    t_last_error = msg;
    return code;
}
int cuda_fail(cudaError_t e, const char* what) {
    t_last_error = std::string(what) + ": " + cudaGetErrorString(e);
    return static_cast<int>(e);
}
void count_launch() { ++t_launch_count; }

// ------------------------------------------------------------------------------------------------
// direct-gather kernel
// ------------------------------------------------------------------------------------------------
struct NoTable {};

template <typename Table>
__device__ __forceinline__ const DevCrop& crop_of(const PreprocParams& P, const Table& T, int z);
template <>
__device__ __forceinline__ const DevCrop& crop_of<NoTable>(const PreprocParams& P, const NoTable&, int z) {
    return P.crops[z];
}
template <>
__device__ __forceinline__ const DevCrop& crop_of<ParamCropTable>(const PreprocParams&, const ParamCropTable& T,
                                                                  int z) {
    return T.c[z];
}

// block = 256 threads = (1 << bw_log2) quads in x  X  (256 >> bw_log2) rows; a quad = 4 output pixels.
// NC = channels of the source pixel, NR = registers per pixel of the chain (4 for a 3-channel source whose chain
// adds an alpha channel, else NC).
template <typename Table, int NC, int NR = NC>
__global__ void __launch_bounds__(256)
preproc_direct_kernel(const __grid_constant__ PreprocParams P, const __grid_constant__ Table T, int bw_log2) {
    const int tid = threadIdx.x;
    const int tx = tid & ((1 << bw_log2) - 1);
    const int ty = tid >> bw_log2;
    const int x0 = ((blockIdx.x << bw_log2) + tx) * 4;
    const int y = blockIdx.y * (256 >> bw_log2) + ty;
    const int z = blockIdx.z;
    if (x0 >= P.W || y >= P.H) return;
    const int nvalid = min(4, P.W - x0);

    float v[4][NC];
    if (z < P.used) {
        gather_quad(P, crop_of<Table>(P, T, z), y, x0, nvalid, v);
    } else {
        fill_background<NC>(P, v);
    }
    if constexpr (NR == NC) {
        apply_program<4, NC>(P.prog, v);
        store_pixels<4, NC>(P, z, y, x0, nvalid, v);
    } else {
        float r[4][NR];
#pragma unroll
        for (int p = 0; p < 4; ++p)
#pragma unroll
            for (int c = 0; c < NR; ++c) r[p][c] = c < NC ? v[p][c] : 0.f;
        apply_program<4, NR>(P.prog, r);
        store_pixels<4, NR>(P, z, y, x0, nvalid, r);
    }
}

// ------------------------------------------------------------------------------------------------
// per-thread context: descriptor staging ring (pinned host + device) for batches that do not fit
// the kernel parameters, and the device buffers of the host-buffer entry point.
// ------------------------------------------------------------------------------------------------
struct Ring {
    static constexpr int kSlots = 8;
    // one slot = [cap tensor maps][cap crop descriptors][3 * cap destination planes], pinned host copy + device copy
    uint8_t* h[kSlots] = {};
    uint8_t* d[kSlots] = {};
    cudaEvent_t ev[kSlots] = {};
    bool pending[kSlots] = {};
    size_t cap = 0;  // crops per slot
    int next = 0;
    int device = -1;
    static size_t slot_bytes(size_t cap) { return cap * (sizeof(CUtensorMap) + sizeof(DevCrop) + 4 * sizeof(DevPlane)); }
    DevPlane* planes_h(int s) const { return reinterpret_cast<DevPlane*>(h[s] + cap * (sizeof(CUtensorMap) + sizeof(DevCrop))); }
    DevPlane* planes_d(int s) const { return reinterpret_cast<DevPlane*>(d[s] + cap * (sizeof(CUtensorMap) + sizeof(DevCrop))); }
    CUtensorMap* maps_h(int s) const { return reinterpret_cast<CUtensorMap*>(h[s]); }
    CUtensorMap* maps_d(int s) const { return reinterpret_cast<CUtensorMap*>(d[s]); }
    DevCrop* crops_h(int s) const { return reinterpret_cast<DevCrop*>(h[s] + cap * sizeof(CUtensorMap)); }
    DevCrop* crops_d(int s) const { return reinterpret_cast<DevCrop*>(d[s] + cap * sizeof(CUtensorMap)); }
};
struct HostPath {
    void* d_img = nullptr;
    size_t img_cap = 0;
    float* d_out = nullptr;
    size_t out_cap = 0;
    int device = -1;
};
constexpr int kHostLanes = 3;  // frames in flight in the host-buffer frame loop (copy engines + SMs overlap)
struct HostLanes {
    cudaStream_t stream[kHostLanes] = {};
    cudaEvent_t done[kHostLanes] = {};
    cudaEvent_t start = nullptr;
    int device = -1;
};
struct Ctx {
    Ring ring;
    HostPath host;           // cvgs_b200_preproc_host: caller's stream
    HostPath lane_buf[kHostLanes];  // cvgs_b200_preproc_host_sequence: one staging set per lane
    HostLanes lanes;
    int sm_count = 0;
    int sm_count_device = -1;
    ~Ctx() {
        // process teardown: the CUDA context may already be gone, so errors are ignored
        for (int i = 0; i < Ring::kSlots; ++i) {
            if (ring.h[i]) cudaFreeHost(ring.h[i]);
            if (ring.d[i]) cudaFree(ring.d[i]);
            if (ring.ev[i]) cudaEventDestroy(ring.ev[i]);
        }
        if (host.d_img) cudaFree(host.d_img);
        if (host.d_out) cudaFree(host.d_out);
        for (int i = 0; i < kHostLanes; ++i) {
            if (lane_buf[i].d_img) cudaFree(lane_buf[i].d_img);
            if (lane_buf[i].d_out) cudaFree(lane_buf[i].d_out);
            if (lanes.stream[i]) cudaStreamDestroy(lanes.stream[i]);
            if (lanes.done[i]) cudaEventDestroy(lanes.done[i]);
        }
        if (lanes.start) cudaEventDestroy(lanes.start);
    }
};
static thread_local Ctx t_ctx;

static int ring_reserve(Ring& r, size_t n, int device) {
    if (r.device == device && n <= r.cap) return CVGS_OK;
    for (int i = 0; i < Ring::kSlots; ++i) {
        if (r.pending[i]) { CVGS_CUDA(cudaEventSynchronize(r.ev[i])); r.pending[i] = false; }
        if (r.h[i]) { CVGS_CUDA(cudaFreeHost(r.h[i])); r.h[i] = nullptr; }
        if (r.d[i]) { CVGS_CUDA(cudaFree(r.d[i])); r.d[i] = nullptr; }
        if (!r.ev[i]) CVGS_CUDA(cudaEventCreateWithFlags(&r.ev[i], cudaEventDisableTiming));
    }
    size_t cap = 256;
    while (cap < n) cap *= 2;
    for (int i = 0; i < Ring::kSlots; ++i) {
        CVGS_CUDA(cudaMallocHost(reinterpret_cast<void**>(&r.h[i]), Ring::slot_bytes(cap)));
        CVGS_CUDA(cudaMalloc(reinterpret_cast<void**>(&r.d[i]), Ring::slot_bytes(cap)));
    }
    r.cap = cap;
    r.device = device;
    return CVGS_OK;
}

int sm_count_of(int device) {
    Ctx& c = t_ctx;
    if (c.sm_count_device != device) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, device) != cudaSuccess || n <= 0) n = 148;
        c.sm_count = n;
        c.sm_count_device = device;
    }
    return c.sm_count;
}

static int pick_bw_log2(int W) {
    const int qw = (W + 3) / 4;
    int best = 3, best_pad = 1 << 30;
    for (int l = 3; l <= 6; ++l) {
        const int bw = 1 << l;
        const int padded = (qw + bw - 1) / bw * bw;
        if (padded <= best_pad) { best_pad = padded; best = l; }
    }
    return best;
}

static int launch_direct(const PreprocParams& P_in, const ParamCropTable* table, cudaStream_t stream) {
    PreprocParams P = P_in;
    specialize_division(P.prog, P.nc, P.bg);  // canonical chains: two-operation division by the launch constants
    const int l = pick_bw_log2(P.W);
    const int qw = (P.W + 3) / 4;
    const int rows = 256 >> l;
    const dim3 grid((qw + (1 << l) - 1) >> l, (P.H + rows - 1) / rows, P.n_planes);
    if (grid.y > 65535u || grid.z > 65535u) return fail(CVGS_ERR_INVALID_VALUE, "batch or height too large for one launch");
    if (P.nc == 3 && P.prog.nregs == 4) {  // cvtColor<BGR2BGRA / BGR2RGBA>: a fourth register per pixel
        if (table) preproc_direct_kernel<ParamCropTable, 3, 4><<<grid, 256, 0, stream>>>(P, *table, l);
        else preproc_direct_kernel<NoTable, 3, 4><<<grid, 256, 0, stream>>>(P, NoTable{}, l);
    } else if (P.nc == 4) {
        if (table) preproc_direct_kernel<ParamCropTable, 4><<<grid, 256, 0, stream>>>(P, *table, l);
        else preproc_direct_kernel<NoTable, 4><<<grid, 256, 0, stream>>>(P, NoTable{}, l);
    } else if (table) {
        preproc_direct_kernel<ParamCropTable, 3><<<grid, 256, 0, stream>>>(P, *table, l);
    } else {
        preproc_direct_kernel<NoTable, 3><<<grid, 256, 0, stream>>>(P, NoTable{}, l);
    }
    count_launch();
    CVGS_CUDA(cudaGetLastError());
    return CVGS_OK;
}

// Image-mode preparation: bind every crop to a cached per-image tensor map.  Returns the number of distinct maps
// written to `maps` (<= max_maps), or -1 when the launch cannot use image maps (unknown / inconsistent parents,
// more maps than fit): the caller then falls back to one map per crop.  On failure the crops are left untouched.
static thread_local ImageMapCache t_image_maps;
static int prepare_image_maps(DevCrop* dc, const cvgs_parent_t* parents, int used, const TmaGeom& G, int W,
                              CUtensorMap* maps, int max_maps, int pb = 3) {
    if (!parents || used <= 0) return -1;
    struct Slot { uintptr_t datastart; int rb; };
    Slot slots[kTmaImageMaps];
    int n_maps = 0;
    if (max_maps > kTmaImageMaps) max_maps = kTmaImageMaps;
    // nothing is modified until every crop is known to fit
    struct Place { int32_t xb, y0, pad; };
    Place place_small[kTmaParamCrops];
    std::vector<Place> place_big;
    Place* place = place_small;
    if (used > kTmaParamCrops) {
        place_big.resize(used);
        place = place_big.data();
    }
    uintptr_t last_ds = 0;
    int last_rb = -1, last_k = -1;
    for (int i = 0; i < used; ++i) {
        const cvgs_parent_t& p = parents[i];
        if (!p.datastart || p.whole_width <= 0 || p.whole_height <= 0) return -1;
        const int rb = rb_class(band_row_bytes(std::min(32 * G.NPB, W), dc[i].fx, pb));
        if (rb == 0 || 4 * rb + kSlotHeader > G.slot_bytes) return -1;
        const uintptr_t ds = reinterpret_cast<uintptr_t>(p.datastart);
        int k = last_k;
        if (ds != last_ds || rb != last_rb) {
            for (k = 0; k < n_maps; ++k)
                if (slots[k].datastart == ds && slots[k].rb == rb) break;
            if (k == n_maps) {
                if (n_maps == max_maps) return -1;
                const CUtensorMap* m = t_image_maps.get(ds, dc[i].pitch, pb * p.whole_width, p.whole_height, rb);
                if (!m) return -1;
                maps[n_maps] = *m;
                slots[n_maps++] = Slot{ds, rb};
            }
            last_ds = ds;
            last_rb = rb;
            last_k = k;
        }
        if (!tma_place_in_image(dc[i], ds, p.whole_width, p.whole_height, rb, k, place[i].xb, place[i].y0, place[i].pad, pb))
            return -1;
    }
    for (int i = 0; i < used; ++i) {
        dc[i].m.xb = place[i].xb;
        dc[i].m.y0 = place[i].y0;
        dc[i].pad = place[i].pad;
    }
    return n_maps;
}

static int ring_reserve(struct Ring& r, size_t n, int device);
// NV12 / NV21 frames through the TMA-staged kernel (preproc_yuv_tma.cuh).  taken = false: the batch is not of the
// shape that kernel is built for and nothing was launched (the caller falls back to the direct-gather kernel).
static int launch_yuv_tma(const cvgs_crop_t* crops, int n_planes, int used, const cvgs_pipeline_t* pipe, const PreprocParams& P,
                          int device, int sms, cudaStream_t stream, bool& taken);

// Shared by the batch entry points and the host-buffer entry point.
// replicas / n_replicas: cvgs_b200_preproc_launch_replicated -- the tensor is written at out and at every replicas[d].
static int preproc_launch_impl(const cvgs_crop_t* crops, const cvgs_parent_t* parents, int n_planes, int used,
                               const cvgs_pipeline_t* pipe, float* out, cudaStream_t stream,
                               void* const* replicas = nullptr, int n_replicas = 0) {
    if (int rc = validate_pipeline(pipe)) return rc;
    if (!out) return fail(CVGS_ERR_INVALID_VALUE, "output pointer is NULL");
    if (n_planes <= 0) return fail(CVGS_ERR_INVALID_VALUE, "n_planes must be positive");
    if (used < 0) return fail(CVGS_ERR_INVALID_VALUE, "used must be non-negative");
    if (used > n_planes) used = n_planes;
    if (used > 0 && !crops) return fail(CVGS_ERR_INVALID_VALUE, "crops is NULL");

    // Frame loops repeat the same pipeline with a new output pointer: the normalised parameters are memoised
    // (thread-local, keyed by the bytes of the pipeline struct without `out`, and the batch size).
    struct ParamMemo {
        bool valid = false;
        cvgs_pipeline_t key;
        int n_planes = 0, used = 0;
        PreprocParams P;
    };
    static thread_local ParamMemo memo;
    PreprocParams P;
    {
        cvgs_pipeline_t key = *pipe;
        key.out = nullptr;
        if (memo.valid && memo.n_planes == n_planes && memo.used == used && std::memcmp(&memo.key, &key, sizeof key) == 0) {
            P = memo.P;
            P.out.base = out;
            P.out.vec4 = out_vec4(P.out, P.W, pipe->out_layout == CVGS_OUT_PLANES, out);
        } else {
            if (int rc = build_params(*pipe, n_planes, used, out, P)) return rc;
            std::memset(&memo.key, 0, sizeof memo.key);
            memo.key = key;
            memo.n_planes = n_planes;
            memo.used = used;
            memo.P = P;
            memo.valid = true;
        }
    }

    int device = 0;
    CVGS_CUDA(cudaGetDevice(&device));
    const int variant = g_variant.load(std::memory_order_relaxed);
    const int sms = sm_count_of(device);

    const bool planes_out = pipe->out_layout == CVGS_OUT_PLANES;  // needs a device table of destinations: ring path
    if (CVGS_IS_YUV(P.src_type) && variant != 1 && n_replicas == 0) {
        bool taken = false;
        if (int rc = launch_yuv_tma(crops, n_planes, used, pipe, P, device, sms, stream, taken)) return rc;
        if (taken) return CVGS_OK;
        if (variant == 2) return fail(CVGS_ERR_NOT_SUPPORTED, "TMA-staged kernel cannot take this input");
    }
    if (n_replicas > 0) {
        const bool fast = !P.band_test && P.used == P.n_planes && P.out.px_stride == 1 && !planes_out && !P.out.u8;
        if (!fast || variant == 1 || n_replicas + 1 > kMaxDest)
            return fail(CVGS_ERR_NOT_SUPPORTED, "replicated output: IGNORE_AR, every plane used, planar float tensors, at most 7 replicas");
    }
    // CV_8UC4 batches the TMA-staged kernel can take go through the 256-crop table or the descriptor ring (the
    // instantiations built for four channels, tma_launch_kernel); everything else of a small batch is decided here
    bool small_batch = used <= kTmaParamCrops && !planes_out && n_replicas == 0;
    // (so do gray / alpha chains on crops without a parent image: their instantiations exist for the image-mode tables and
    // the descriptor ring, not for one-map-per-crop parameter tables)
    if (small_batch && variant != 1 && (P.src_type != CVGS_8UC3 || (!parents && (gray_program(P) || alpha_program(P))))) {
        DevCrop probe[kTmaParamCrops];
        bool ok = true;
        for (int i = 0; i < used && ok; ++i) ok = fill_crop(crops[i], *pipe, i, probe[i]) == CVGS_OK;
        TmaGeom g;
        if (ok && tma_plan(P, probe, used, n_planes, sms, parents != nullptr, 1, g)) small_batch = false;
    }
    if (small_batch) {
        // small batch: descriptors (and tensor maps) ride in the kernel parameters -- no staging copy,
        // graph-capturable
        alignas(64) TmaParamTable tt;  // also serves as the image-mode table (its first maps / same crop array offset
                                       // are copied into a TmaImageTable below)
        const double t0 = now_us();
        for (int i = 0; i < used; ++i)
            if (int rc = fill_crop(crops[i], *pipe, i, tt.c[i])) return rc;
        const double t1 = now_us();
        TmaParams K;
        K.P = P;
        K.maps = nullptr;
        if (variant != 1 && P.src_type == CVGS_8UC3 && tma_plan(P, tt.c, used, n_planes, sms, parents != nullptr, items_per_warp(), K.G)) {
            const int chain = scaled_program_for(P, K);
            const int pb = 3;
            const MemRange src = crops_range(tt.c, used, pb);  // before the TMA fields overwrite the pointers
            const double t2 = now_us();
            // image mode first (cached maps, small parameter block); else one map per crop
            alignas(64) TmaImageTable it;
            DevCrop saved[kTmaParamCrops];
            std::memcpy(saved, tt.c, static_cast<size_t>(used) * sizeof(DevCrop));
            const int n_img = pb == 3 ? prepare_image_maps(tt.c, parents, used, K.G, P.W, it.m, kTmaImageMaps) : -1;
            int rc = -1;
            if (n_img >= 0) {
                std::memcpy(it.c, tt.c, static_cast<size_t>(used) * sizeof(DevCrop));
                const double t3 = now_us();
                K.G.pdl_wait = overlap_needs_wait(stream, out_range(P), src) ? 1 : 0;
                rc = tma_launch_kernel<TmaImageTable>(K, it, chain, device, stream);
                const double t4 = now_us();
                t_prof.fill += t1 - t0; t_prof.plan += t2 - t1; t_prof.encode += t3 - t2; t_prof.launch += t4 - t3;
                ++t_prof.calls;
                return rc;
            }
            bool ok = true;
            // (one map per crop: the gray instantiation is not built for that table -- such batches take the direct kernel)
            for (int i = 0; i < used && ok; ++i) ok = pb == 3 && !chain_needs_image_table(chain) && tma_prepare_crop(tt.c[i], K.G, P.W, i, &tt.m[i]) == CVGS_OK;
            const double t3 = now_us();
            if (ok) {
                K.G.pdl_wait = overlap_needs_wait(stream, out_range(P), src) ? 1 : 0;
                rc = tma_launch_kernel<TmaParamTable>(K, tt, chain, device, stream);
                const double t4 = now_us();
                t_prof.fill += t1 - t0; t_prof.plan += t2 - t1; t_prof.encode += t3 - t2; t_prof.launch += t4 - t3;
                ++t_prof.calls;
                return rc;
            }
            // the driver refused a tensor map (exotic geometry): the direct-gather kernel takes anything
            std::memcpy(tt.c, saved, static_cast<size_t>(used) * sizeof(DevCrop));
        }
        if (variant == 2) return fail(CVGS_ERR_NOT_SUPPORTED, "TMA-staged kernel cannot take this input");
        ParamCropTable table;
        std::memcpy(table.c, tt.c, static_cast<size_t>(used) * sizeof(DevCrop));
        overlap_forget(stream);
        return launch_direct(P, &table, stream);
    }

    if (used <= kTmaImageCrops && parents && variant != 1 && !planes_out && n_replicas == 0) {
        // medium batch with named parent images: a few cached maps + the descriptors still fit the kernel
        // parameters (13 KB), so there is no staging copy in front of the kernel
        alignas(64) TmaImageTableL lt;
        bool filled = true;
        for (int i = 0; i < used && filled; ++i) filled = fill_crop(crops[i], *pipe, i, lt.c[i]) == CVGS_OK;
        if (!filled) return CVGS_ERR_INVALID_VALUE;  // message already recorded by fill_crop
        TmaParams K;
        K.P = P;
        K.maps = nullptr;
        if (tma_plan(P, lt.c, used, n_planes, sms, true, items_per_warp(), K.G)) {
            const int chain = scaled_program_for(P, K);
            const int pb = pixel_bytes_of(P.src_type);
            const MemRange src = crops_range(lt.c, used, pb);
            if (prepare_image_maps(lt.c, parents, used, K.G, P.W, lt.m, kTmaImageMaps, pb) >= 0) {
                K.G.pdl_wait = overlap_needs_wait(stream, out_range(P), src) ? 1 : 0;
                return tma_launch_kernel<TmaImageTableL>(K, lt, chain, device, stream);
            }
        }
    }

    Ring& r = t_ctx.ring;
    if (int rc = ring_reserve(r, static_cast<size_t>(std::max(used, n_planes)), device)) return rc;
    const int slot = r.next;
    r.next = (r.next + 1) % Ring::kSlots;
    if (r.pending[slot]) { CVGS_CUDA(cudaEventSynchronize(r.ev[slot])); r.pending[slot] = false; }
    DevCrop* hc = r.crops_h(slot);
    for (int i = 0; i < used; ++i)
        if (int rc = fill_crop(crops[i], *pipe, i, hc[i])) return rc;
    TmaParams K;
    K.P = P;
    bool use_tma = variant != 1 && tma_plan(P, hc, used, n_planes, sms, parents != nullptr, 1, K.G);
    int chain = CH_GENERIC;
    size_t map_bytes = 0;
    if (use_tma) {
        chain = scaled_program_for(P, K);
        if (chain_needs_image_table(chain) && n_replicas > 0)  // no peer-store instantiation of these chains
            return fail(CVGS_ERR_NOT_SUPPORTED, "replicated launches take chains that keep the channel count");
        CUtensorMap* hm = r.maps_h(slot);
        // maps sit in front of the crops in the slot; image mode needs only a few of them
        const int pb = pixel_bytes_of(P.src_type);
        const int n_img = prepare_image_maps(hc, parents, used, K.G, P.W, hm, kTmaImageMaps, pb);
        if (n_img >= 0) {
            map_bytes = static_cast<size_t>(n_img) * sizeof(CUtensorMap);
        } else {
            for (int i = 0; i < used && use_tma; ++i) use_tma = tma_prepare_crop(hc[i], K.G, P.W, i, &hm[i], pb) == CVGS_OK;
            map_bytes = static_cast<size_t>(used) * sizeof(CUtensorMap);
            if (!use_tma)  // restore the pointers the failed preparation overwrote
                for (int i = 0; i < used; ++i)
                    if (int rc = fill_crop(crops[i], *pipe, i, hc[i])) return rc;
        }
    }
    if (!use_tma && (variant == 2 || n_replicas > 0)) return fail(CVGS_ERR_NOT_SUPPORTED, "TMA-staged kernel cannot take this input");
    if (planes_out) {
        // destination images, indexed [z][source channel] on the device (the channel reorder is applied here)
        const cvgs_plane_t* hp = static_cast<const cvgs_plane_t*>(pipe->out);
        DevPlane* dp = r.planes_h(slot);
        const int nc = P.nc;
        for (int z = 0; z < n_planes; ++z)
            for (int c = 0; c < nc; ++c) {
                const cvgs_plane_t& q = hp[z * nc + P.prog.dst_chan[c]];
                if (!q.data || q.pitch_bytes < 4LL * P.W || (q.pitch_bytes & 3))
                    return fail(CVGS_ERR_INVALID_VALUE, "plane " + std::to_string(z) + ": bad destination image");
                dp[z * nc + c].data = static_cast<float*>(q.data);
                dp[z * nc + c].pitch = q.pitch_bytes / 4;
            }
        CVGS_CUDA(cudaMemcpyAsync(r.planes_d(slot), dp, static_cast<size_t>(n_planes) * nc * sizeof(DevPlane),
                                  cudaMemcpyHostToDevice, stream));
        P.out.planes = r.planes_d(slot);
        K.P.out.planes = r.planes_d(slot);
    }
    // one or two copies: the maps actually used, and the crop descriptors
    if (use_tma)
        CVGS_CUDA(cudaMemcpyAsync(r.maps_d(slot), r.maps_h(slot), map_bytes, cudaMemcpyHostToDevice, stream));
    CVGS_CUDA(cudaMemcpyAsync(r.crops_d(slot), hc, static_cast<size_t>(used) * sizeof(DevCrop), cudaMemcpyHostToDevice, stream));
    int rc;
    if (use_tma) {
        K.P.crops = r.crops_d(slot);
        K.maps = r.maps_d(slot);
        // the descriptor copy sits between this kernel and its predecessor: plain stream order, no overlap
        K.G.pdl_wait = 1;
        overlap_forget(stream);
        if (n_replicas > 0) {
            K.n_dest = n_replicas + 1;
            K.dest_delta[0] = 0;
            for (int d = 0; d < n_replicas; ++d) {
                const long long diff = reinterpret_cast<const char*>(replicas[d]) - reinterpret_cast<const char*>(out);
                if (!replicas[d] || (diff & 3)) return fail(CVGS_ERR_INVALID_VALUE, "replica " + std::to_string(d) + ": NULL or not float-aligned");
                K.dest_delta[d + 1] = diff / 4;
            }
            rc = tma_launch_replicated(K, chain, device, stream);
        } else {
            rc = tma_launch_kernel<TmaNoTable>(K, TmaNoTable{0}, chain, device, stream);
        }
    } else {
        P.crops = r.crops_d(slot);
        overlap_forget(stream);
        rc = launch_direct(P, nullptr, stream);
    }
    CVGS_CUDA(cudaEventRecord(r.ev[slot], stream));
    r.pending[slot] = true;
    return rc;
}

// ------------------------------------------------------------------------------------------------
// Coalesced frame groups: the crops of several independent argument sets (own parent frame, own output tensor, same
// pipeline) in ONE launch of the TMA-staged kernel.  A 50-crop frame is 7.5 MB of traffic -- 1.5 us of HBM time, less
// than a kernel launch costs the host and the GPU front end -- so a frame loop that launches per frame is bound by
// launch granularity; ten frames per launch are not.  Used by cvgs_b200_preproc_launch_sequence_ex.
// ------------------------------------------------------------------------------------------------
static thread_local DevMapCache t_dev_maps;

static int launch_yuv_tma(const cvgs_crop_t* crops, int n_planes, int used, const cvgs_pipeline_t* pipe, const PreprocParams& P,
                          int device, int sms, cudaStream_t stream, bool& taken) {
    taken = false;
    if (used != n_planes || used <= 0) return CVGS_OK;
    Ring& r = t_ctx.ring;
    if (int rc = ring_reserve(r, static_cast<size_t>(n_planes), device)) return rc;
    const int slot = r.next;
    std::vector<DevCrop> dc(static_cast<size_t>(used));
    for (int i = 0; i < used; ++i)
        if (int rc = fill_crop(crops[i], *pipe, i, dc[i])) return rc;
    YuvParams K;
    K.P = P;
    if (!yuv_tma_plan(P, dc.data(), used, n_planes, sms, K.G)) return CVGS_OK;
    DevMapCache& mc = t_dev_maps;
    if (int rc = mc.reserve(device)) return rc;
    if (r.pending[slot]) { CVGS_CUDA(cudaEventSynchronize(r.ev[slot])); r.pending[slot] = false; }
    static_assert(sizeof(DevYuv) == sizeof(DevCrop), "the frame table lives in the ring's crop region");
    DevYuv* hf = reinterpret_cast<DevYuv*>(r.crops_h(slot));
    const int TWp = std::min(32 * K.G.NPB, P.W);
    const bool packed = P.src_type == CVGS_Y210;
    const int depth = packed ? 2 : yuv_depth_of(P.src_type);
    const int csh = P.src_type == CVGS_P210 ? 0 : 1;
    for (int attempt = 0;; ++attempt) {
        const uint32_t gen = mc.generation;
        bool restart = false;
        for (int i = 0; i < used && !restart; ++i) {
            const DevCrop& c = dc[i];
            DevYuv& f = hf[i];
            const uintptr_t luma = reinterpret_cast<uintptr_t>(c.data);
            const uintptr_t chroma = luma + static_cast<uintptr_t>(c.pitch) * static_cast<uintptr_t>(c.h);
            f.xbL = static_cast<int32_t>(luma & 15);
            f.xbC = static_cast<int32_t>(chroma & 15);
            f.w = c.w;
            f.h = c.h;
            f.fx = c.fx;
            f.fy = c.fy;
            f.rbL = packed ? yuv_rb_packed(TWp, c.fx) : yuv_rb_luma(TWp, c.fx, depth);
            f.rbC = packed ? 0 : yuv_rb_chroma(TWp, c.fx, depth);
            f.pad0 = f.pad1 = 0;
            f.mapL = mc.get(luma, c.pitch, (packed ? 4 : depth) * c.w, c.h, f.rbL);  // luma plane: w samples x h rows (Y210: w / 2 groups)
            f.mapC = f.mapL < 0 ? -1 : packed ? f.mapL : mc.get(chroma, c.pitch, depth * c.w, c.h >> csh, f.rbC);  // chroma: w / 2 pairs x h / 2 (4:2:0) or h rows
            if (f.mapL < 0 || f.mapC < 0) return CVGS_OK;  // the driver refused the geometry: direct-gather kernel
            if (mc.generation != gen) restart = true;      // the table started over: indices handed out so far are void
        }
        if (!restart) break;
        if (attempt > 0) return CVGS_OK;
    }
    if (int rc = mc.flush()) return rc;
    // chain for interpolated values (2^33 folded into its first op), division constants
    TmaParams tmp;
    tmp.P = P;
    tmp.G = K.G;
    const int chain = scaled_program(P, tmp);
    K.prog_img = tmp.prog_img;
    std::memcpy(K.zh, tmp.zh, sizeof K.zh);
    std::memcpy(K.zl, tmp.zl, sizeof K.zl);
    K.G.explicit_prescale = tmp.G.explicit_prescale;
    if (depth == 1) {
        for (int i = 0; i < 9; ++i) K.m[i] = std::ldexp(P.yuv[i], 100);
        K.yoff = std::ldexp(P.yuv[9], -133);
        K.coff = std::ldexp(P.yuv[10], -133);
        const uint32_t kU = P.src_type == CVGS_NV12 ? 0u : 1u;
        K.selU1 = 0x4044u | kU << 8;
        K.selV1 = 0x4044u | (kU ^ 1u) << 8;
    } else {  // sample * 2^-135 in, (RGB * 64) * 2^-33 out (P.yuv[11] = 64 is in the exponent)
        for (int i = 0; i < 9; ++i) K.m[i] = std::ldexp(P.yuv[i], 108);
        K.yoff = std::ldexp(P.yuv[9], -135);
        K.coff = std::ldexp(P.yuv[10], -135);
        K.selU1 = 0x4104u;
        K.selV1 = 0x4324u;
    }
    K.csh = csh;
    K.maps = mc.d;
    K.frames = reinterpret_cast<const DevYuv*>(r.crops_d(slot));
    r.next = (r.next + 1) % Ring::kSlots;
    CVGS_CUDA(cudaMemcpyAsync(r.crops_d(slot), hf, static_cast<size_t>(used) * sizeof(DevYuv), cudaMemcpyHostToDevice, stream));
    overlap_forget(stream);
    const int rc = packed       ? (chain == CH_FMA_DIV ? yuv_launch_instance<CH_FMA_DIV, 2, true>(K, device, stream) : yuv_launch_instance<CH_GENERIC, 2, true>(K, device, stream))
                   : depth == 1 ? (chain == CH_FMA_DIV ? yuv_launch_instance<CH_FMA_DIV>(K, device, stream) : yuv_launch_instance<CH_GENERIC>(K, device, stream))
                                : (chain == CH_FMA_DIV ? yuv_launch_instance<CH_FMA_DIV, 2>(K, device, stream) : yuv_launch_instance<CH_GENERIC, 2>(K, device, stream));
    CVGS_CUDA(cudaEventRecord(r.ev[slot], stream));
    r.pending[slot] = true;
    taken = rc == CVGS_OK;
    return rc;
}

struct MultiSet {
    const cvgs_crop_t* crops;
    const cvgs_parent_t* parents;
    int n;
    float* out;
};
constexpr int kMultiDeclined = -2;  // not an error: the caller launches the sets one by one

static int launch_multi(const MultiSet* sets, int G, const cvgs_pipeline_t* pipe, bool early_wait, cudaStream_t stream) {
    int total = 0;
    for (int g = 0; g < G; ++g) total += sets[g].n;
    if (G <= 0 || G > kMultiGroups || total <= 0 || total > kMultiCrops) return kMultiDeclined;
    PreprocParams P;
    if (int rc = build_params(*pipe, total, total, sets[0].out, P)) return rc;
    int device = 0;
    CVGS_CUDA(cudaGetDevice(&device));
    const int sms = sm_count_of(device);
    DevMapCache& mc = t_dev_maps;
    if (int rc = mc.reserve(device)) return rc;
    alignas(64) TmaMultiTable mt;
    static thread_local std::vector<DevCrop> full_tl;  // the crops in full form (planning, placement); 45 KB
    full_tl.resize(kMultiCrops);
    DevCrop* full = full_tl.data();
    TmaParams K;
    for (int attempt = 0;; ++attempt) {
        int z = 0;
        for (int g = 0; g < G; ++g) {
            mt.out_base[g] = sets[g].out;
            mt.z_first[g] = z;
            for (int i = 0; i < sets[g].n; ++i, ++z)
                if (int rc = fill_crop(sets[g].crops[i], *pipe, z, full[z])) return rc;
        }
        K.P = P;
        K.P.out.base = nullptr;  // every plane goes through out_base[]
        static const int grid_div = [] {  // tuning override (profiling)
            const char* e = std::getenv("CVGS_B200_MULTI_GRID_DIV");
            return e ? std::max(1, std::atoi(e)) : 1;
        }();
        if (!tma_plan(P, full, total, total, sms, true, 1, K.G, true, grid_div)) return kMultiDeclined;
        const uint32_t gen = mc.generation;
        const int TWp = std::min(32 * K.G.NPB, P.W);
        bool restart = false;
        z = 0;
        for (int g = 0; g < G && !restart; ++g) {
            uintptr_t last_ds = 0;
            int last_rb = -1, last_idx = -1;
            for (int i = 0; i < sets[g].n; ++i, ++z) {
                const cvgs_parent_t& p = sets[g].parents[i];
                const DevCrop& c = full[z];
                if (!p.datastart || p.whole_width <= 0 || p.whole_height <= 0) return kMultiDeclined;
                const int rb = rb_class(band_row_bytes(TWp, c.fx), true);
                if (rb == 0 || 4 * rb + kSlotHeader > K.G.slot_bytes) return kMultiDeclined;
                const uintptr_t ds = reinterpret_cast<uintptr_t>(p.datastart);
                int idx = last_idx;
                if (ds != last_ds || rb != last_rb) {
                    idx = mc.get(ds, c.pitch, 3 * p.whole_width, p.whole_height, rb);
                    if (idx < 0) return kMultiDeclined;
                    if (mc.generation != gen) {  // the table started over: indices handed out so far are void
                        restart = true;
                        break;
                    }
                    last_ds = ds;
                    last_rb = rb;
                    last_idx = idx;
                }
                DevCropC& o = mt.c[z];
                if (!tma_place_in_image(c, ds, p.whole_width, p.whole_height, rb, idx, o.xb, o.y0, o.pad)) return kMultiDeclined;
                o.w = c.w;
                o.h = c.h;
                o.fx = c.fx;
                o.fy = c.fy;
                o.group = g;
            }
        }
        if (!restart) break;
        if (attempt > 0) return kMultiDeclined;
    }
    if (int rc = mc.flush()) return rc;
    const int chain = scaled_program(P, K);
    K.P.crops = nullptr;
    K.maps = mc.d;
    K.G.pdl_wait = early_wait ? 1 : 0;
    return tma_launch_multi(K, mt, chain, device, stream);
}

// ------------------------------------------------------------------------------------------------
// Host-buffer path: upload of the part of a frame its crops touch.  Default: one cudaMemcpy2DAsync of the rows [min y,
// max y + h).  Optional (cvgs_b200_set_host_upload(1)), when the caller's frame is pinned host memory that the device can
// read in place: a small kernel pulls the 128-byte x 16-row tiles that some crop overlaps into the device staging
// image -- 4.4 MB instead of 6.0 MB on the bench's 1080p frames (the exact union of the rectangles is 3.8 MB).  Measured
// on the PCIe Gen5 boxes of this pool it LOSES: SM-issued reads of host memory are 128-byte requests and reach 29 GB/s
// where the copy engine streams 47 GB/s, so 332 K against 386 K crops/s end to end; it is kept for hosts whose GPUs read
// host memory at link rate (NVLink-C2C).  The tile mask rides in the kernel parameters.
// ------------------------------------------------------------------------------------------------
constexpr int kTileBytes = 128;        // tile width (one PCIe read request per row of a tile)
constexpr int kTileMaskWords = 1008;   // 32256 tiles: ~4 KB of kernel parameters
struct TileMask {
    uint32_t bits[kTileMaskWords];
};
__global__ void __launch_bounds__(256) upload_tiles_kernel(const uint8_t* __restrict__ host_img, long long host_pitch,
                                                           uint8_t* __restrict__ dev_img, long long dev_pitch, int row_chunks,
                                                           int height, int tiles_x, int tile_rows, int n_tiles,
                                                           const __grid_constant__ TileMask mask) {
    const int tile = blockIdx.x * 8 + (threadIdx.x >> 5);  // one warp per tile
    if (tile >= n_tiles || !((mask.bits[tile >> 5] >> (tile & 31)) & 1u)) return;
    const int lane = threadIdx.x & 31;
    const int ty = tile / tiles_x, tx = tile - ty * tiles_x;
    const int chunk = tx * (kTileBytes / 16) + (lane & 7);  // 16-byte chunk of the row
    if (chunk >= row_chunks) return;
    const int y0 = ty * tile_rows, y1 = min(height, y0 + tile_rows);
    // every lane keeps up to four 16-byte reads in flight (rows y, y + 4, y + 8, y + 12): PCIe latency is microseconds
    for (int y = y0 + (lane >> 3); y < y1; y += 16) {
        uint4 v[4];
#pragma unroll
        for (int k = 0; k < 4; ++k)
            if (y + 4 * k < y1) v[k] = *reinterpret_cast<const uint4*>(host_img + (long long)(y + 4 * k) * host_pitch + 16LL * chunk);
#pragma unroll
        for (int k = 0; k < 4; ++k)
            if (y + 4 * k < y1) *reinterpret_cast<uint4*>(dev_img + (long long)(y + 4 * k) * dev_pitch + 16LL * chunk) = v[k];
    }
}

// Device-visible address of a pinned host buffer, or nullptr (pageable memory, or the query failed).  Frame buffers
// recur: a small per-thread cache keeps cudaPointerGetAttributes off the per-frame path.
static const uint8_t* pinned_device_view(const void* host_ptr) {
    struct Entry { const void* host; const uint8_t* dev; };
    static thread_local Entry cache[64] = {};
    Entry& e = cache[(reinterpret_cast<uintptr_t>(host_ptr) >> 12) & 63];
    if (e.host == host_ptr) return e.dev;
    cudaPointerAttributes a;
    const uint8_t* dev = nullptr;
    if (cudaPointerGetAttributes(&a, host_ptr) == cudaSuccess && a.type == cudaMemoryTypeHost && a.devicePointer)
        dev = static_cast<const uint8_t*>(a.devicePointer);
    else
        (void)cudaGetLastError();
    e.host = host_ptr;
    e.dev = dev;
    return dev;
}
static thread_local uint64_t t_h2d_bytes = 0, t_d2h_bytes = 0;  // cvgs_b200_debug_host_bytes

static int host_reserve(HostPath& h, size_t img_bytes, size_t out_bytes, int device) {
    if (h.device != device) {  // the old buffers belong to the old device: free them there
        if (h.device >= 0 && (h.d_img || h.d_out)) {
            CVGS_CUDA(cudaSetDevice(h.device));
            if (h.d_img) cudaFree(h.d_img);
            if (h.d_out) cudaFree(h.d_out);
            CVGS_CUDA(cudaSetDevice(device));
        }
        h.img_cap = h.out_cap = 0;
        h.d_img = nullptr;
        h.d_out = nullptr;
        h.device = device;
    }
    if (img_bytes > h.img_cap) {
        if (h.d_img) { CVGS_CUDA(cudaDeviceSynchronize()); CVGS_CUDA(cudaFree(h.d_img)); }
        CVGS_CUDA(cudaMalloc(&h.d_img, img_bytes));
        h.img_cap = img_bytes;
    }
    if (out_bytes > h.out_cap) {
        if (h.d_out) { CVGS_CUDA(cudaDeviceSynchronize()); CVGS_CUDA(cudaFree(h.d_out)); }
        CVGS_CUDA(cudaMalloc(reinterpret_cast<void**>(&h.d_out), out_bytes));
        h.out_cap = out_bytes;
    }
    return CVGS_OK;
}

}  // namespace cvgs

// ------------------------------------------------------------------------------------------------
// C-ABI
// ------------------------------------------------------------------------------------------------
using namespace cvgs;

// Helper threads of the frame loop live for the rest of the process (their per-thread caches -- tensor maps, memoised
// programs -- stay warm between sequences, and a sequence does not pay for thread creation).  The pool is allocated
// once and never destroyed: at process exit the helpers are parked on the condition variable and simply end with the
// process.
struct SeqPool {
    std::mutex mu;
    std::condition_variable cv_job, cv_done;
    std::atomic<uint64_t> generation{0};
    std::function<void(int)> job;  // called with the worker index 1 .. job_workers - 1
    int job_workers = 0;
    int pending = 0;
    int n_threads = 0;             // helpers started so far (worker indices 1 .. n_threads)
    void helper_main(int w) {
        uint64_t seen = 0;
        for (;;) {
            // a short spin first: frame loops arrive back to back, and a futex wake costs tens of microseconds
            for (int i = 0; i < 20000 && generation.load(std::memory_order_acquire) == seen; ++i) {
#if defined(__x86_64__)
                __builtin_ia32_pause();
#endif
            }
            std::unique_lock<std::mutex> lk(mu);
            cv_job.wait(lk, [&] { return generation.load(std::memory_order_acquire) != seen; });
            seen = generation.load(std::memory_order_acquire);
            if (w >= job_workers) continue;
            std::function<void(int)> f = job;
            lk.unlock();
            f(w);
            lk.lock();
            if (--pending == 0) cv_done.notify_all();
        }
    }
    // runs f(1) .. f(workers - 1) on the helpers; returns after the caller-supplied `own` ran on this thread and every
    // helper finished
    template <typename Own>
    void run(int workers, std::function<void(int)> f, Own own) {
        {
            std::lock_guard<std::mutex> lk(mu);
            while (n_threads < workers - 1) {
                const int w = ++n_threads;
                std::thread([this, w] { helper_main(w); }).detach();
            }
            job = std::move(f);
            job_workers = workers;
            pending = workers - 1;
            generation.fetch_add(1, std::memory_order_release);
        }
        cv_job.notify_all();
        own();
        std::unique_lock<std::mutex> lk(mu);
        cv_done.wait(lk, [&] { return pending == 0; });
        job = nullptr;
    }
};
static SeqPool* seq_pool() {
    // intentionally leaked, see above.  A forked child inherits the pool's bookkeeping but none of its threads: it gets
    // a pool of its own (callers hold g_seq_mu, so this is not racy).
    static SeqPool* p = nullptr;
    static pid_t owner = 0;
    if (!p || owner != getpid()) {
        p = new SeqPool;
        owner = getpid();
    }
    return p;
}

extern "C" {

int cvgs_b200_version(void) { return CVGS_B200_VERSION; }
const char* cvgs_b200_last_error(void) { return t_last_error.c_str(); }
int cvgs_b200_set_kernel_variant(int variant) { return g_variant.exchange(variant); }
int cvgs_b200_set_overlap(int mode) { return g_overlap.exchange(mode < 0 ? 0 : (mode > 2 ? 2 : mode)); }
int cvgs_b200_set_coalesce(int enable) { return g_coalesce.exchange(enable ? 1 : 0); }
int cvgs_b200_set_host_upload(int mode) { return g_host_tiles.exchange(mode ? 1 : 0); }
int64_t cvgs_b200_launch_count(void) { return t_launch_count; }
// Diagnostics for the CPU test-suite: the launch plan of the TMA kernel for a batch geometry (no device needed).
// out[0..11] = {ok, NPB, HP, tiles_x, total_items, slot_bytes, slots, resident, grid, max row bytes needed,
//               items covered by the per-warp ranges, 1 if the ranges tile [0, total_items) in order without gaps}
int cvgs_b200_debug_program(const cvgs_pipeline_t* pipeline, float* out80) {
    if (!pipeline || !out80) return fail(CVGS_ERR_INVALID_VALUE, "NULL argument");
    if (int rc = validate_pipeline(pipeline)) return rc;
    DevProgram prog;
    if (int rc = build_program(*pipeline, prog)) return rc;
    for (int i = 0; i < 80; ++i) out80[i] = 0.f;
    out80[0] = static_cast<float>(prog.n_ops);
    out80[1] = static_cast<float>(prog.nc_out);
    out80[2] = static_cast<float>(prog.nregs);
    out80[3] = static_cast<float>(prog.special);
    for (int r = 0; r < 4; ++r) out80[4 + r] = static_cast<float>(prog.dst_chan[r]);
    for (int i = 0; i < prog.n_ops && i < 8; ++i) {
        out80[8 + 9 * i] = static_cast<float>(prog.ops[i].kind);
        for (int c = 0; c < 4; ++c) {
            out80[8 + 9 * i + 1 + c] = prog.ops[i].a[c];
            out80[8 + 9 * i + 5 + c] = prog.ops[i].b[c];
        }
    }
    return CVGS_OK;
}
int cvgs_b200_debug_plan(const cvgs_crop_t* crops, int32_t n_planes, int32_t used, const cvgs_pipeline_t* pipeline,
                         int32_t sm_count, int32_t image_mode, int32_t items_per_warp_, int64_t* out12) {
    if (!pipeline || !out12) return fail(CVGS_ERR_INVALID_VALUE, "NULL argument");
    if (int rc = validate_pipeline(pipeline)) return rc;
    for (int i = 0; i < 12; ++i) out12[i] = 0;
    PreprocParams P;
    static float dummy_out;
    if (int rc = build_params(*pipeline, n_planes, used, &dummy_out, P)) return rc;
    std::vector<DevCrop> dc(static_cast<size_t>(std::max(used, 1)));
    int rb_need = 0;
    for (int i = 0; i < used; ++i)
        if (int rc = fill_crop(crops[i], *pipeline, i, dc[i])) return rc;
    TmaGeom G{};
    if (!tma_plan(P, dc.data(), used, n_planes, sm_count, image_mode != 0, std::max(1, items_per_warp_), G, false)) return CVGS_OK;
    for (int i = 0; i < used; ++i) {
        int rb = band_row_bytes(std::min(32 * G.NPB, P.W), dc[i].fx);
        if (image_mode) rb = rb_class(rb);
        rb_need = std::max(rb_need, rb);
    }
    long long covered = 0, expect_first = 0;
    bool tiled = true;
    for (int g = 0; g < G.grid * kWarps; ++g) {
        long long first, count;
        host_item_range(G, g, first, count);
        if (first != expect_first || count < 0) tiled = false;
        expect_first = first + count;
        covered += count;
    }
    const int64_t vals[12] = {1, G.NPB, G.HP, G.tiles_x, G.total_items, G.slot_bytes, G.slots, G.resident, G.grid, rb_need,
                              covered, tiled && expect_first == G.total_items ? 1 : 0};
    for (int i = 0; i < 12; ++i) out12[i] = vals[i];
    return CVGS_OK;
}

// Diagnostics for the CPU test-suite: the overlap bookkeeping (pure host logic).  Returns 1 when a launch with these
// output / source byte ranges on `stream_key` would have to wait for its predecessor, 0 when it may overlap.
int cvgs_b200_debug_overlap_query(void* stream_key, uint64_t out_lo, uint64_t out_hi, uint64_t src_lo, uint64_t src_hi) {
    MemRange o, s;
    o.lo = static_cast<uintptr_t>(out_lo); o.hi = static_cast<uintptr_t>(out_hi);
    s.lo = static_cast<uintptr_t>(src_lo); s.hi = static_cast<uintptr_t>(src_hi);
    return overlap_needs_wait(static_cast<cudaStream_t>(stream_key), o, s) ? 1 : 0;
}

static bool warp_denominators_normal(const float* m, int w_padded, int h);
int cvgs_b200_debug_warp_mode(const float* m9, int32_t warp_type, int32_t dst_width, int32_t dst_height) {
    if (!m9 || dst_width <= 0 || dst_height <= 0) return -1;
    if (warp_type == CVGS_WARP_AFFINE) return WM_AFFINE;
    return warp_denominators_normal(m9, (dst_width + 127) / 128 * 128, dst_height) ? WM_PERSPECTIVE_NORMAL : WM_PERSPECTIVE;
}

int cvgs_b200_debug_chain_kind(const cvgs_pipeline_t* pipeline, float* alpha) {
    if (int rc = validate_pipeline(pipeline)) return -rc;
    float dummy = 0.f;
    PreprocParams P;
    if (int rc = build_params(*pipeline, 1, 1, &dummy, P)) return -rc;
    TmaParams K;
    K.P = P;
    std::memset(&K.G, 0, sizeof K.G);
    const int pb = pixel_bytes_of(P.src_type);
    K.G.prescale = pb >= 6 ? kPreScale16 : kPreScale;
    const int chain = scaled_program_for(P, K);
    if (alpha) *alpha = K.alpha;
    return chain;
}

uint32_t cvgs_b200_debug_fast_div(uint32_t n, uint32_t d) {
    return d ? fast_div(n, fast_div_make(d)) : 0u;
}

int cvgs_b200_debug_host_bytes(uint64_t* h2d, uint64_t* d2h, int reset) {
    if (h2d) *h2d = t_h2d_bytes;
    if (d2h) *d2h = t_d2h_bytes;
    if (reset) t_h2d_bytes = t_d2h_bytes = 0;
    return CVGS_OK;
}

int cvgs_b200_debug_host_profile(double* out5, int reset) {
    if (out5) {
        out5[0] = static_cast<double>(t_prof.calls);
        out5[1] = t_prof.fill; out5[2] = t_prof.plan; out5[3] = t_prof.encode; out5[4] = t_prof.launch;
    }
    if (reset) t_prof = HostProfile{};
    return CVGS_OK;
}

int cvgs_b200_preproc_launch(const cvgs_crop_t* crops, int32_t n_planes, int32_t used,
                             const cvgs_pipeline_t* pipeline, void* stream) {
    CVGS_RANGE("cvgs_b200_preproc_launch");
    if (!pipeline) return fail(CVGS_ERR_INVALID_VALUE, "pipeline is NULL");
    return preproc_launch_impl(crops, nullptr, n_planes, used, pipeline, static_cast<float*>(pipeline->out),
                               static_cast<cudaStream_t>(stream));
}

int cvgs_b200_preproc_launch_ex(const cvgs_crop_t* crops, const cvgs_parent_t* parents, int32_t n_planes, int32_t used,
                                const cvgs_pipeline_t* pipeline, void* stream) {
    CVGS_RANGE("cvgs_b200_preproc_launch_ex");
    if (!pipeline) return fail(CVGS_ERR_INVALID_VALUE, "pipeline is NULL");
    return preproc_launch_impl(crops, parents, n_planes, used, pipeline, static_cast<float*>(pipeline->out),
                               static_cast<cudaStream_t>(stream));
}

int cvgs_b200_preproc_launch_rects(const void* frame, int32_t frame_width, int32_t frame_height, int32_t frame_pitch,
                                   const cvgs_rect_t* rects, int32_t n_planes, int32_t used, const cvgs_pipeline_t* pipeline,
                                   void* stream) {
    CVGS_RANGE("cvgs_b200_preproc_launch_rects");
    if (!pipeline) return fail(CVGS_ERR_INVALID_VALUE, "pipeline is NULL");
    if (!frame || frame_width <= 0 || frame_height <= 0 || frame_pitch <= 0) return fail(CVGS_ERR_INVALID_VALUE, "bad frame");
    if (used < 0 || n_planes <= 0) return fail(CVGS_ERR_INVALID_VALUE, "bad batch size");
    if (used > n_planes) used = n_planes;
    if (used > 0 && !rects) return fail(CVGS_ERR_INVALID_VALUE, "rects is NULL");
    if (CVGS_IS_YUV(pipeline->src_type)) return fail(CVGS_ERR_NOT_SUPPORTED, "YUV frames are read whole: no rectangles");
    const int px = pixel_bytes_of(pipeline->src_type);
    if (static_cast<long long>(px) * frame_width > frame_pitch && frame_height > 1) return fail(CVGS_ERR_INVALID_VALUE, "frame pitch smaller than a row");
    // fk::Crop: thread (x, y) of plane i reads (x + rect.x, y + rect.y) of the frame, rect.width x rect.height threads
    // (crop.cuh:23-55) -- the ROI pointer form of the batch launch, with the frame named as the parent of every crop
    std::vector<cvgs_crop_t> crops(static_cast<size_t>(used));
    std::vector<cvgs_parent_t> parents(static_cast<size_t>(used), cvgs_parent_t{frame, frame_width, frame_height});
    for (int i = 0; i < used; ++i) {
        const cvgs_rect_t& r = rects[i];
        if (r.x < 0 || r.y < 0 || r.width <= 0 || r.height <= 0 || r.x > frame_width - r.width || r.y > frame_height - r.height)
            return fail(CVGS_ERR_INVALID_VALUE, "rect " + std::to_string(i) + " outside the frame");
        crops[i].data = static_cast<const uint8_t*>(frame) + static_cast<size_t>(r.y) * frame_pitch + static_cast<size_t>(px) * r.x;
        crops[i].width = r.width;
        crops[i].height = r.height;
        crops[i].pitch = frame_pitch;
        crops[i].reserved = 0;
    }
    return preproc_launch_impl(crops.data(), pipeline->src_type == CVGS_8UC3 ? parents.data() : nullptr, n_planes, used, pipeline,
                               static_cast<float*>(pipeline->out), static_cast<cudaStream_t>(stream));
}

int cvgs_b200_preproc_launch_replicated(const cvgs_crop_t* crops, const cvgs_parent_t* parents, int32_t n_planes, int32_t used,
                                        const cvgs_pipeline_t* pipeline, void* const* replicas, int32_t n_replicas,
                                        void* stream) {
    CVGS_RANGE("cvgs_b200_preproc_launch_replicated");
    if (!pipeline) return fail(CVGS_ERR_INVALID_VALUE, "pipeline is NULL");
    if (n_replicas < 0 || (n_replicas > 0 && !replicas)) return fail(CVGS_ERR_INVALID_VALUE, "bad replica list");
    return preproc_launch_impl(crops, parents, n_planes, used, pipeline, static_cast<float*>(pipeline->out),
                               static_cast<cudaStream_t>(stream), replicas, n_replicas);
}

// Peer-mapped device memory for the replicated launch: thin wrappers over cudaMalloc / cudaIpc* so that a host language
// without CUDA bindings (and torch-free callers) can place one tensor per GPU and map the others' into its own process.
int cvgs_b200_dev_alloc(void** ptr, uint64_t bytes) {
    if (!ptr) return fail(CVGS_ERR_INVALID_VALUE, "ptr is NULL");
    CVGS_CUDA(cudaMalloc(ptr, static_cast<size_t>(bytes)));
    return CVGS_OK;
}
int cvgs_b200_dev_free(void* ptr) {
    CVGS_CUDA(cudaFree(ptr));
    return CVGS_OK;
}
int cvgs_b200_ipc_export(void* ptr, void* handle64) {
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
    if (!ptr || !handle64) return fail(CVGS_ERR_INVALID_VALUE, "NULL argument");
    cudaIpcMemHandle_t h;
    CVGS_CUDA(cudaIpcGetMemHandle(&h, ptr));
    std::memcpy(handle64, &h, sizeof h);
    return CVGS_OK;
}
int cvgs_b200_ipc_open(const void* handle64, void** ptr) {
    if (!ptr || !handle64) return fail(CVGS_ERR_INVALID_VALUE, "NULL argument");
    cudaIpcMemHandle_t h;
    std::memcpy(&h, handle64, sizeof h);
    CVGS_CUDA(cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return CVGS_OK;
}
int cvgs_b200_ipc_close(void* ptr) {
    CVGS_CUDA(cudaIpcCloseMemHandle(ptr));
    return CVGS_OK;
}

// Perspective: the denominator m6 x + m7 y + m8 is linear in the destination pixel, so over the (padded) destination
// rectangle its extremes sit at the corners.  True when every float denominator the kernel computes is finite, of one sign
// and of magnitude in [2^-100, 2^100] -- the range where the correctly rounded reciprocal needs no special cases (the fast
// kernel then skips the range check of __frcp_rn).  The float evaluation FADD(FFMA(m6, x, FMUL(m7, y)), m8) is within
// 3 * 2^-24 * S of the exact value, S = |m6| xmax + |m7| ymax + |m8|; the margin below is 2^-20 * S.
static bool warp_denominators_normal(const float* m, int w_padded, int h) {
    const double a = m[6], b = m[7], c = m[8];
    if (!std::isfinite(a) || !std::isfinite(b) || !std::isfinite(c)) return false;
    const double xs[2] = {0.0, static_cast<double>(w_padded - 1)}, ys[2] = {0.0, static_cast<double>(h - 1)};
    const double S = std::fabs(a) * xs[1] + std::fabs(b) * ys[1] + std::fabs(c);
    double lo = INFINITY, hi = 0.0;
    int sign = 0;
    for (double x : xs)
        for (double y : ys) {
            const double e = a * x + b * y + c;  // exact enough: doubles carry 53 bits, the products of floats and ints < 2^24 are exact
            const int sg = e > 0 ? 1 : (e < 0 ? -1 : 0);
            if (sg == 0 || (sign != 0 && sg != sign)) return false;
            sign = sg;
            lo = std::min(lo, std::fabs(e));
            hi = std::max(hi, std::fabs(e));
        }
    return lo >= S * 9.5367431640625e-07 && lo >= 7.888609052210118e-31 && hi <= 1.2676506002282294e+30;
}

// Batched warp: descriptors ride in the kernel parameters, kWarpParamPlanes planes per launch.
static int warp_launch_impl(const cvgs_crop_t* images, const cvgs_warp_t* warps, int n_planes, int used,
                            const cvgs_pipeline_t* pipe, cudaStream_t stream) {
    if (int rc = validate_pipeline(pipe)) return rc;
    if (CVGS_IS_YUV(pipe->src_type)) return fail(CVGS_ERR_NOT_SUPPORTED, "warp takes CV_8U / CV_16U / CV_16S sources with 3 or 4 channels");
    if (!pipe->out) return fail(CVGS_ERR_INVALID_VALUE, "output pointer is NULL");
    if (n_planes <= 0) return fail(CVGS_ERR_INVALID_VALUE, "n_planes must be positive");
    if (used < 0) return fail(CVGS_ERR_INVALID_VALUE, "used must be non-negative");
    if (used > n_planes) used = n_planes;
    if (used > 0 && (!images || !warps)) return fail(CVGS_ERR_INVALID_VALUE, "images / warps is NULL");
    cvgs_pipeline_t q = *pipe;
    q.interp_mode = CVGS_INTERP_FLOAT;  // fk::Warping hands the chain the float interpolation
    q.aspect_mode = CVGS_IGNORE_AR;
    PreprocParams P;
    if (int rc = build_params(q, n_planes, used, static_cast<float*>(pipe->out), P)) return rc;
    if (P.prog.special) return fail(CVGS_ERR_NOT_SUPPORTED, "warp: conversions that change the channel count are not on this path");
    specialize_division(P.prog, P.nc, P.bg);  // pixels outside the source image are 0, unused planes the background
    for (int i = 0; i < used; ++i) {
        DevCrop scratch;
        if (int rc = fill_crop(images[i], q, i, scratch)) return rc;
        if (warps[i].type != CVGS_WARP_AFFINE && warps[i].type != CVGS_WARP_PERSPECTIVE)
            return fail(CVGS_ERR_INVALID_VALUE, "warp " + std::to_string(i) + ": bad type");
    }
    int device = 0;
    CVGS_CUDA(cudaGetDevice(&device));
    overlap_forget(stream);
    if (pipe->out_layout == CVGS_OUT_PLANES) {
        Ring& r = t_ctx.ring;
        if (int rc = ring_reserve(r, static_cast<size_t>(n_planes), device)) return rc;
        const int slot = r.next;
        r.next = (r.next + 1) % Ring::kSlots;
        if (r.pending[slot]) { CVGS_CUDA(cudaEventSynchronize(r.ev[slot])); r.pending[slot] = false; }
        const cvgs_plane_t* hp = static_cast<const cvgs_plane_t*>(pipe->out);
        DevPlane* dp = r.planes_h(slot);
        const int nc = P.nc;
        for (int z = 0; z < n_planes; ++z)
            for (int c = 0; c < nc; ++c) {
                const cvgs_plane_t& pl = hp[z * nc + P.prog.dst_chan[c]];
                if (!pl.data || pl.pitch_bytes < 4LL * P.W || (pl.pitch_bytes & 3))
                    return fail(CVGS_ERR_INVALID_VALUE, "plane " + std::to_string(z) + ": bad destination image");
                dp[z * nc + c].data = static_cast<float*>(pl.data);
                dp[z * nc + c].pitch = pl.pitch_bytes / 4;
            }
        CVGS_CUDA(cudaMemcpyAsync(r.planes_d(slot), dp, static_cast<size_t>(n_planes) * nc * sizeof(DevPlane),
                                  cudaMemcpyHostToDevice, stream));
        P.out.planes = r.planes_d(slot);
        CVGS_CUDA(cudaEventRecord(r.ev[slot], stream));  // re-recorded after the kernels below
        r.pending[slot] = true;
    }
    // fast instantiation (preproc_warp.cuh): float tensor rows, chain = [MUL | FMA | ADD] [two-operation DIV]
    WarpFastParams F{};
    int fast = 0;  // 1: linear chain, 2: linear chain + division
    {
        const DevProgram& g = P.prog;
        const bool lin0 = g.n_ops >= 1 && (g.ops[0].kind == DOP_MUL || g.ops[0].kind == DOP_FMA || g.ops[0].kind == DOP_ADD);
        const bool shape = g.n_ops == 0 || (g.n_ops == 1 && (lin0 || g.ops[0].kind == DOP_DIVC)) ||
                           (g.n_ops == 2 && lin0 && g.ops[1].kind == DOP_DIVC);
        bool ok = shape && !g.round_u8 && !P.out.u8 && !P.out.planes && P.out.px_stride == 1 && g_variant.load(std::memory_order_relaxed) != 1;  // variant 1: the general kernel (tests)
        // offsets inside a destination plane are 32-bit in the kernel
        ok = ok && std::llabs(P.out.row_stride) * P.H + P.W < 0x1fffffffLL;
        for (int c = 0; c < P.nc && ok; ++c) ok = std::llabs(P.out.c_off[c]) < (1LL << 40);
        for (int i = 0; i < used && ok; ++i)  // float(width), float(height) exact
            ok = images[i].width <= (1 << 24) && images[i].height <= (1 << 24);
        if (ok) {
            const bool has_div = g.n_ops > 0 && g.ops[g.n_ops - 1].kind == DOP_DIVC;
            fast = has_div ? 2 : 1;
            for (int c = 0; c < 4; ++c) {
                F.a[c] = lin0 && g.ops[0].kind != DOP_ADD ? g.ops[0].a[c] : 1.0f;  // MUL(a) == FMA(a, -0), ADD(b) == FMA(1, b)
                F.b[c] = !lin0 || g.ops[0].kind == DOP_MUL ? -0.0f : (g.ops[0].kind == DOP_ADD ? g.ops[0].a[c] : g.ops[0].b[c]);
                F.zh[c] = has_div ? g.ops[g.n_ops - 1].a[c] : 1.0f;
                F.zl[c] = has_div ? g.ops[g.n_ops - 1].b[c] : 0.0f;
                F.bg[c] = P.bg[c];
                F.c_off[c] = c < P.nc ? P.out.c_off[c] : 0;
            }
            F.W = P.W;
            F.H = P.H;
            F.used = used;
            F.base = P.out.base;
            F.z_stride = P.out.z_stride;
            F.row_stride = P.out.row_stride;
        }
    }
    const dim3 block(256);
    for (int z0 = 0; z0 < n_planes; z0 += kWarpParamPlanes) {
        const int nz = std::min(kWarpParamPlanes, n_planes - z0);
        alignas(64) WarpTable T;
        for (int i = 0; i < nz && z0 + i < used; ++i) {
            const cvgs_crop_t& c = images[z0 + i];
            DevWarp& d = T.w[i];
            d.data = static_cast<const uint8_t*>(c.data);
            d.w = c.width;
            d.h = c.height;
            d.pitch = c.pitch;
            d.type = warps[z0 + i].type;
            std::memcpy(d.m, warps[z0 + i].m, sizeof d.m);
            d.pad = d.type == CVGS_WARP_AFFINE ? WM_AFFINE : warp_denominators_normal(d.m, (P.W + 127) / 128 * 128, P.H) ? WM_PERSPECTIVE_NORMAL : WM_PERSPECTIVE;
        }
        // four pixels per thread measured best from 640x640 planes up and within 10% below (1, 2 and 4 tried on B200)
        const dim3 grid((P.W + 127) / 128, (P.H + 7) / 8, nz);
        if (fast) {
            // two rows per warp halve the per-thread set-up per pixel; worth it once the launch still fills the GPU several times
            const long long ctas1 = static_cast<long long>(grid.x) * grid.y * nz;
            F.rows = ctas1 >= 8LL * 5 * sm_count_of(device) ? 2 : 1;
            if (const char* e = std::getenv("CVGS_WARP_ROWS")) F.rows = std::max(1, std::atoi(e));  // tuning override (profiling)
            const dim3 grid(((P.W + 127) / 128), (P.H + 8 * F.rows - 1) / (8 * F.rows), nz);
#define CVGS_WARP_FAST(T_, NC_)                                                                         \
    if (fast == 2) preproc_warp_fast_kernel<T_, NC_, true><<<grid, block, 0, stream>>>(F, T, z0);       \
    else preproc_warp_fast_kernel<T_, NC_, false><<<grid, block, 0, stream>>>(F, T, z0)
            switch (pipe->src_type) {
                case CVGS_8UC3: CVGS_WARP_FAST(unsigned char, 3); break;
                case CVGS_8UC4: CVGS_WARP_FAST(unsigned char, 4); break;
                case CVGS_16UC3: CVGS_WARP_FAST(unsigned short, 3); break;
                case CVGS_16UC4: CVGS_WARP_FAST(unsigned short, 4); break;
                case CVGS_16SC3: CVGS_WARP_FAST(short, 3); break;
                default: CVGS_WARP_FAST(short, 4); break;
            }
#undef CVGS_WARP_FAST
            CVGS_CUDA(cudaGetLastError());
            ++t_launch_count;
            continue;
        }
        switch (pipe->src_type) {
            case CVGS_8UC3: preproc_warp_kernel<4, unsigned char, 3><<<grid, block, 0, stream>>>(P, T, z0); break;
            case CVGS_8UC4: preproc_warp_kernel<4, unsigned char, 4><<<grid, block, 0, stream>>>(P, T, z0); break;
            case CVGS_16UC3: preproc_warp_kernel<4, unsigned short, 3><<<grid, block, 0, stream>>>(P, T, z0); break;
            case CVGS_16UC4: preproc_warp_kernel<4, unsigned short, 4><<<grid, block, 0, stream>>>(P, T, z0); break;
            case CVGS_16SC3: preproc_warp_kernel<4, short, 3><<<grid, block, 0, stream>>>(P, T, z0); break;
            default: preproc_warp_kernel<4, short, 4><<<grid, block, 0, stream>>>(P, T, z0); break;
        }
        CVGS_CUDA(cudaGetLastError());
        ++t_launch_count;
    }
    if (pipe->out_layout == CVGS_OUT_PLANES) {
        Ring& r = t_ctx.ring;
        const int slot = (r.next + Ring::kSlots - 1) % Ring::kSlots;
        CVGS_CUDA(cudaEventRecord(r.ev[slot], stream));
    }
    return CVGS_OK;
}

int cvgs_b200_warp_launch(const cvgs_crop_t* images, const cvgs_warp_t* warps, int32_t n_planes, int32_t used,
                          const cvgs_pipeline_t* pipeline, void* stream) {
    CVGS_RANGE("cvgs_b200_warp_launch");
    if (!pipeline) return fail(CVGS_ERR_INVALID_VALUE, "pipeline is NULL");
    return warp_launch_impl(images, warps, n_planes, used, pipeline, static_cast<cudaStream_t>(stream));
}

static int preproc_host_impl(HostPath& h, const void* host_image, int32_t image_width, int32_t image_height,
                             int32_t image_pitch, const cvgs_rect_t* rects, int32_t n_planes, int32_t used,
                             const cvgs_pipeline_t* pipeline, float* host_out, void* stream_,
                             cudaEvent_t download_after = nullptr, cudaEvent_t download_done = nullptr) {
    if (int rc = validate_pipeline(pipeline)) return rc;
    if (!host_image || !host_out || !rects) return fail(CVGS_ERR_INVALID_VALUE, "NULL host buffer");
    if (pipeline->src_type != CVGS_8UC3) return fail(CVGS_ERR_NOT_SUPPORTED, "the host-buffer entry point takes CV_8UC3 frames");
    if (pipeline->dst_type == CVGS_8UC3 || pipeline->dst_type == CVGS_8UC4 || pipeline->out_row_pitch != 0 ||
        pipeline->out_layout == CVGS_OUT_PLANES)
        return fail(CVGS_ERR_NOT_SUPPORTED, "the host-buffer entry point writes one float tensor");
    if (image_width <= 0 || image_height <= 0 || image_pitch < 3 * image_width)
        return fail(CVGS_ERR_INVALID_VALUE, "bad host image geometry");
    if (n_planes <= 0 || used < 0) return fail(CVGS_ERR_INVALID_VALUE, "bad batch size");
    if (used > n_planes) used = n_planes;
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    int device = 0;
    CVGS_CUDA(cudaGetDevice(&device));
    // device copy keeps rows 512-byte aligned like cudaMallocPitch would
    const size_t d_pitch = (static_cast<size_t>(3) * image_width + 511) / 512 * 512;
    const size_t img_bytes = d_pitch * image_height;
    // the chain may change the channel count (cvtColor to BGRA / gray): the tensor has what it produces
    DevProgram prog;
    if (int rc = build_program(*pipeline, prog)) return rc;
    const size_t out_floats = static_cast<size_t>(prog.nc_out) * pipeline->dst_width * pipeline->dst_height * n_planes;
    if (int rc = host_reserve(h, img_bytes, out_floats * sizeof(float), device)) return rc;

    // upload only the rows some crop touches
    int y_lo = image_height, y_hi = 0;
    std::vector<cvgs_crop_t> crops(static_cast<size_t>(used));
    for (int i = 0; i < used; ++i) {
        const cvgs_rect_t& r = rects[i];
        if (r.x < 0 || r.y < 0 || r.width <= 0 || r.height <= 0 || r.x + r.width > image_width ||
            r.y + r.height > image_height)
            return fail(CVGS_ERR_INVALID_VALUE, "rect " + std::to_string(i) + " outside the image");
        y_lo = std::min(y_lo, r.y);
        y_hi = std::max(y_hi, r.y + r.height);
        crops[i].data = static_cast<const uint8_t*>(h.d_img) + static_cast<size_t>(r.y) * d_pitch + 3 * static_cast<size_t>(r.x);
        crops[i].width = r.width;
        crops[i].height = r.height;
        crops[i].pitch = static_cast<int32_t>(d_pitch);
        crops[i].reserved = 0;
    }
    if (used > 0) {
        const uint8_t* pinned = g_host_tiles.load(std::memory_order_relaxed) ? pinned_device_view(host_image) : nullptr;
        const int row_chunks = (3 * image_width + 15) / 16;
        const int tiles_x = (3 * image_width + kTileBytes - 1) / kTileBytes;
        int tile_rows = 16;
        while (static_cast<long long>(tiles_x) * ((image_height + tile_rows - 1) / tile_rows) > 32LL * kTileMaskWords) tile_rows *= 2;
        const int tiles_y = (image_height + tile_rows - 1) / tile_rows;
        if (pinned && (reinterpret_cast<uintptr_t>(pinned) & 15) == 0 && (image_pitch & 15) == 0 &&
            16LL * row_chunks <= image_pitch) {
            alignas(64) TileMask mask;
            const int n_tiles = tiles_x * tiles_y;
            std::memset(mask.bits, 0, static_cast<size_t>((n_tiles + 31) / 32) * 4);
            long long tiles_set = 0;
            for (int i = 0; i < used; ++i) {
                const cvgs_rect_t& r = rects[i];
                const int tx0 = 3 * r.x / kTileBytes, tx1 = (3 * (r.x + r.width) - 1) / kTileBytes;
                const int ty0 = r.y / tile_rows, ty1 = (r.y + r.height - 1) / tile_rows;
                for (int ty = ty0; ty <= ty1; ++ty)
                    for (int t = ty * tiles_x + tx0; t <= ty * tiles_x + tx1; ++t) {
                        uint32_t& w = mask.bits[t >> 5];
                        const uint32_t b = 1u << (t & 31);
                        tiles_set += !(w & b);
                        w |= b;
                    }
            }
            upload_tiles_kernel<<<(n_tiles + 7) / 8, 256, 0, stream>>>(pinned, image_pitch, static_cast<uint8_t*>(h.d_img),
                                                                      static_cast<long long>(d_pitch), row_chunks, image_height,
                                                                      tiles_x, tile_rows, n_tiles, mask);
            CVGS_CUDA(cudaGetLastError());
            count_launch();
            overlap_forget(stream);
            t_h2d_bytes += static_cast<uint64_t>(tiles_set) * kTileBytes * tile_rows;  // upper bound: edge tiles are clipped
        } else {
            CVGS_CUDA(cudaMemcpy2DAsync(static_cast<uint8_t*>(h.d_img) + static_cast<size_t>(y_lo) * d_pitch, d_pitch,
                                        static_cast<const uint8_t*>(host_image) + static_cast<size_t>(y_lo) * image_pitch,
                                        static_cast<size_t>(image_pitch), static_cast<size_t>(3) * image_width,
                                        static_cast<size_t>(y_hi - y_lo), cudaMemcpyHostToDevice, stream));
            t_h2d_bytes += static_cast<uint64_t>(3) * image_width * (y_hi - y_lo);
        }
    }
    cvgs_pipeline_t p = *pipeline;
    p.out_plane_stride = 0;
    // every crop is a rectangle of the staging image: per-image tensor maps
    std::vector<cvgs_parent_t> parents(static_cast<size_t>(used), cvgs_parent_t{h.d_img, image_width, image_height});
    if (int rc = preproc_launch_impl(crops.data(), parents.data(), n_planes, used, &p, h.d_out, stream)) return rc;
    // frame loop: downloads complete in frame order whatever lane they run on (a later frame may write the same
    // host tensor as an earlier one)
    if (download_after) CVGS_CUDA(cudaStreamWaitEvent(stream, download_after, 0));
    CVGS_CUDA(cudaMemcpyAsync(host_out, h.d_out, out_floats * sizeof(float), cudaMemcpyDeviceToHost, stream));
    t_d2h_bytes += out_floats * sizeof(float);
    if (download_done) CVGS_CUDA(cudaEventRecord(download_done, stream));
    return CVGS_OK;
}

int cvgs_b200_preproc_host(const void* host_image, int32_t image_width, int32_t image_height, int32_t image_pitch,
                           const cvgs_rect_t* rects, int32_t n_planes, int32_t used,
                           const cvgs_pipeline_t* pipeline, float* host_out, void* stream) {
    CVGS_RANGE("cvgs_b200_preproc_host");
    return preproc_host_impl(t_ctx.host, host_image, image_width, image_height, image_pitch, rects, n_planes, used, pipeline,
                             host_out, stream);
}

int cvgs_b200_preproc_launch_sequence(const cvgs_crop_t* const* crops, const int32_t* n_planes, const int32_t* used,
                                      const cvgs_pipeline_t* const* pipelines, int32_t n_sets, int32_t steps,
                                      void* stream) {
    CVGS_RANGE("cvgs_b200_preproc_launch_sequence");
    if (!crops || !n_planes || !used || !pipelines || n_sets <= 0 || steps < 0)
        return fail(CVGS_ERR_INVALID_VALUE, "bad sequence arguments");
    for (int i = 0; i < steps; ++i) {
        const int s = i % n_sets;
        if (int rc = cvgs_b200_preproc_launch(crops[s], n_planes[s], used[s], pipelines[s], stream)) return rc;
    }
    return CVGS_OK;
}

// ------------------------------------------------------------------------------------------------
// Frame loop over several host threads.  One launch costs the host 3-4 us (descriptor fill, plan, tensor-map lookup,
// cudaLaunchKernelEx) and the GPU about as long for a 50-crop frame, so a loop driven by one thread is host-bound.
// When the argument sets of a sequence are provably independent (outputs pairwise disjoint, no output overlaps a
// source) the loop is split by argument set over the calling thread and a few helper threads, each launching into
// its own stream: set s always goes to worker s % T, so repeated uses of one set stay in stream order, and the
// caller's stream is forked into / joined from the helper streams with events -- seen from that stream the loop is
// still one operation.  Only with cvgs_b200_set_overlap(1) (the caller allowed the library to reorder independent
// launches) and outside stream capture.
// ------------------------------------------------------------------------------------------------
constexpr int kSeqMaxWorkers = 4;
struct SeqStreams {
    cudaStream_t stream[kSeqMaxWorkers] = {};
    cudaEvent_t done[kSeqMaxWorkers] = {};
    cudaEvent_t start = nullptr;
    int device = -1;
};
static std::mutex g_seq_mu;  // one multi-threaded sequence at a time per process (they would share the helper streams)
static SeqStreams g_seq;

static int seq_workers() {
    static const int n = [] {
        const char* e = std::getenv("CVGS_B200_SEQ_THREADS");
        return e ? std::max(1, std::min(kSeqMaxWorkers, std::atoi(e))) : 3;  // measured: 1: 3.85, 2: 2.73, 3: 2.32, 4: 2.29 us per frame
    }();
    return n;
}

// Conservative independence proof over the argument sets of a sequence.
static bool sequence_sets_independent(const cvgs_crop_t* const* crops, const int32_t* n_planes, const int32_t* used,
                                      const cvgs_pipeline_t* const* pipelines, int n_sets) {
    if (n_sets > 256) return false;
    std::vector<MemRange> outs(n_sets), srcs(n_sets);
    for (int s = 0; s < n_sets; ++s) {
        const cvgs_pipeline_t* p = pipelines[s];
        if (!p || !p->out || !crops[s] || n_planes[s] <= 0 || used[s] < 0) return false;
        if (validate_pipeline(p) != CVGS_OK) return false;
        if (p->out_layout == CVGS_OUT_PLANES || p->dst_type == CVGS_8UC3 || p->dst_type == CVGS_8UC4 || CVGS_IS_YUV(p->src_type)) return false;
        PreprocParams P;
        if (build_params(*p, n_planes[s], std::min(used[s], n_planes[s]), static_cast<float*>(p->out), P) != CVGS_OK) return false;
        const long long plane = static_cast<long long>(P.W) * P.H;
        const int nc = P.prog.nc_out;
        const long long extent = P.out.px_stride == 1 ? (nc - 1) * P.out.c_stride + (P.n_planes - 1) * P.out.z_stride + plane
                                                      : (P.n_planes - 1) * P.out.z_stride + P.out.row_stride * P.H;
        outs[s].lo = reinterpret_cast<uintptr_t>(p->out);
        outs[s].hi = outs[s].lo + static_cast<uintptr_t>(extent) * sizeof(float);
        const int px = pixel_bytes_of(p->src_type);
        MemRange r;
        r.lo = ~static_cast<uintptr_t>(0);
        for (int i = 0; i < std::min(used[s], n_planes[s]); ++i) {
            const cvgs_crop_t& c = crops[s][i];
            if (!c.data || c.width <= 0 || c.height <= 0 || c.pitch <= 0) return false;
            const uintptr_t lo = reinterpret_cast<uintptr_t>(c.data);
            r.lo = std::min(r.lo, lo);
            r.hi = std::max(r.hi, lo + static_cast<uintptr_t>(c.height - 1) * c.pitch + static_cast<uintptr_t>(px) * c.width);
        }
        if (r.lo > r.hi) r.lo = r.hi = 0;
        srcs[s] = r;
    }
    for (int a = 0; a < n_sets; ++a)
        for (int b = 0; b < n_sets; ++b) {
            if (outs[a].overlaps(srcs[b])) return false;
            if (a < b && outs[a].overlaps(outs[b])) return false;
        }
    return true;
}

// Can the argument sets ride in shared launches (launch_multi)?  Same pipeline in every set apart from the output
// pointer, the common geometry (CV_8UC3, IGNORE_AR, every plane used, tight or strided NCHW float), parents named.
static bool sequence_sets_coalescible(const cvgs_parent_t* const* parents, const int32_t* n_planes, const int32_t* used,
                                      const cvgs_pipeline_t* const* pipelines, int n_sets) {
    if (!g_coalesce.load(std::memory_order_relaxed) || n_sets < 2) return false;
    const cvgs_pipeline_t& p0 = *pipelines[0];
    if (p0.src_type != CVGS_8UC3 || p0.aspect_mode != CVGS_IGNORE_AR || p0.out_layout != CVGS_OUT_NCHW || p0.out_row_pitch != 0 ||
        (p0.dst_type != 0 && p0.dst_type != CVGS_32FC3))
        return false;
    DevProgram prog;
    if (build_program(p0, prog) != CVGS_OK || prog.special) return false;
    cvgs_pipeline_t k0 = p0;
    k0.out = nullptr;
    for (int s = 0; s < n_sets; ++s) {
        if (!parents[s] || n_planes[s] != used[s] || n_planes[s] <= 0 || n_planes[s] > kMultiCrops) return false;
        cvgs_pipeline_t k = *pipelines[s];
        k.out = nullptr;
        if (std::memcmp(&k, &k0, sizeof k) != 0) return false;
    }
    return true;
}

// The coalesced frame loop: consecutive steps (distinct, independent argument sets) share launches of up to
// kMultiCrops crops.  One host thread, the caller's stream.  Consecutive launches overlap (late griddepcontrol.wait)
// unless a launch rewrites the tensor of a set that a launch possibly still in flight wrote: everything since the
// last launch that waited up front may be in flight, and at most StreamTrack::kWindow launches are allowed to be.
static int sequence_coalesced(const cvgs_crop_t* const* crops, const cvgs_parent_t* const* parents, const int32_t* n_planes,
                              const cvgs_pipeline_t* const* pipelines, int n_sets, int steps, cudaStream_t stream) {
    std::vector<int> last_writer(static_cast<size_t>(n_sets), -1);
    int last_early = 0, L = 0;
    bool force_early = true;  // whatever precedes the sequence on the stream is unknown
    MultiSet ms[kMultiGroups];
    static const int max_groups = [] {  // tuning overrides (profiling)
        const char* e = std::getenv("CVGS_B200_MULTI_GROUPS");
        return e ? std::max(1, std::min(kMultiGroups, std::atoi(e))) : kMultiGroups;
    }();
    static const int max_crops = [] {
        const char* e = std::getenv("CVGS_B200_MULTI_CROPS");
        return e ? std::max(1, std::min(kMultiCrops, std::atoi(e))) : kMultiCrops;
    }();
    overlap_forget(stream);
    for (int i = 0; i < steps;) {
        int G = 0, total = 0;
        bool hazard = force_early || L - last_early >= StreamTrack::kWindow;
        while (i + G < steps && G < max_groups && G < n_sets) {
            const int s = (i + G) % n_sets;
            if (G > 0 && total + n_planes[s] > max_crops) break;
            ms[G] = MultiSet{crops[s], parents[s], n_planes[s], static_cast<float*>(pipelines[s]->out)};
            if (last_writer[static_cast<size_t>(s)] >= last_early) hazard = true;
            total += n_planes[s];
            ++G;
        }
        int rc = launch_multi(ms, G, pipelines[i % n_sets], hazard, stream);
        if (rc == kMultiDeclined) {  // a geometry the shared launch does not take: these sets go one by one, in plain order
            for (int g = 0; g < G; ++g) {
                const int s = (i + g) % n_sets;
                if (int r2 = cvgs_b200_preproc_launch_ex(crops[s], parents[s], n_planes[s], n_planes[s], pipelines[s], stream)) return r2;
            }
            overlap_forget(stream);
            force_early = true;
        } else if (rc != CVGS_OK) {
            return rc;
        } else {
            force_early = false;
            if (hazard) last_early = L;
        }
        for (int g = 0; g < G; ++g) last_writer[static_cast<size_t>((i + g) % n_sets)] = L;
        ++L;
        i += G;
    }
    overlap_forget(stream);
    return CVGS_OK;
}

int cvgs_b200_preproc_launch_sequence_ex(const cvgs_crop_t* const* crops, const cvgs_parent_t* const* parents,
                                         const int32_t* n_planes, const int32_t* used,
                                         const cvgs_pipeline_t* const* pipelines, int32_t n_sets, int32_t steps,
                                         void* stream_) {
    CVGS_RANGE("cvgs_b200_preproc_launch_sequence_ex");
    if (!crops || !parents || !n_planes || !used || !pipelines || n_sets <= 0 || steps < 0)
        return fail(CVGS_ERR_INVALID_VALUE, "bad sequence arguments");
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    // Sets whose single launch fills the GPU several times over (c3: 256 crops -> 224x224 is 57 K items for 2 960 warps) gain
    // nothing from sharing launches or from helper threads -- measured 46.9 us per 256-crop set in shared launches against
    // 39.7-40.9 us launched per set from one thread with the early wait dropped (two launch threads: 39.7-40.5 us once their
    // streams exist, 7 ms to set them up; plain stream order: 42.2 us).
    static const bool big_rule = [] { const char* e = std::getenv("CVGS_B200_SEQ_BIG"); return !(e && e[0] == '0'); }();  // tuning override
    bool big_sets = big_rule;
    if (big_sets) {
        int device = 0;
        if (cudaGetDevice(&device) != cudaSuccess) device = 0;
        const long long fill = 4LL * sm_count_of(device) * kMaxResident * kWarps;
        for (int s = 0; s < n_sets && big_sets; ++s)
            big_sets = pipelines[s] && static_cast<long long>(used[s]) * ((pipelines[s]->dst_height + 1) / 2) *
                                               ((pipelines[s]->dst_width + 127) / 128) >= fill;
    }
    if (!big_sets && steps >= 2 && g_overlap.load(std::memory_order_relaxed) && g_variant.load(std::memory_order_relaxed) != 1) {
        cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
        if (cudaStreamIsCapturing(stream, &cap) == cudaSuccess && cap == cudaStreamCaptureStatusNone &&
            sequence_sets_independent(crops, n_planes, used, pipelines, n_sets) &&
            sequence_sets_coalescible(parents, n_planes, used, pipelines, n_sets))
            return sequence_coalesced(crops, parents, n_planes, pipelines, n_sets, steps, stream);
    }
    int workers = std::min(seq_workers(), static_cast<int>(n_sets));
    if (workers > 1 && (big_sets || steps < 16 || !g_overlap.load(std::memory_order_relaxed))) workers = 1;
    if (workers > 1) {
        cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
        if (cudaStreamIsCapturing(stream, &cap) != cudaSuccess || cap != cudaStreamCaptureStatusNone) workers = 1;
    }
    if (workers > 1 && !sequence_sets_independent(crops, n_planes, used, pipelines, n_sets)) workers = 1;
    std::unique_lock<std::mutex> seq_lock(g_seq_mu, std::defer_lock);
    if (workers > 1 && !seq_lock.try_lock()) workers = 1;
    if (workers <= 1) {
        // one thread, the caller's stream: still a segment the library controls from its first launch to its last
        // (the first launch waits up front: what precedes the loop is unknown)
        overlap_forget(stream);
        const bool may_overlap = g_overlap.load(std::memory_order_relaxed) != 0;
        t_in_frame_loop = may_overlap;
        if (may_overlap) t_items_hint = 4;  // fewer, longer-lived CTAs per launch leave room for the next launch (see items_per_warp)
        int rc = CVGS_OK;
        for (int i = 0; i < steps && rc == CVGS_OK; ++i) {
            const int s = i % n_sets;
            rc = cvgs_b200_preproc_launch_ex(crops[s], parents[s], n_planes[s], used[s], pipelines[s], stream);
        }
        t_in_frame_loop = false;
        t_items_hint = 0;
        overlap_forget(stream);
        return rc;
    }

    int device = 0;
    CVGS_CUDA(cudaGetDevice(&device));
    SeqStreams& Q = g_seq;
    if (Q.device != device) {
        for (int i = 1; i < kSeqMaxWorkers; ++i) {
            if (Q.stream[i]) { cudaStreamDestroy(Q.stream[i]); Q.stream[i] = nullptr; }
            if (Q.done[i]) { cudaEventDestroy(Q.done[i]); Q.done[i] = nullptr; }
            CVGS_CUDA(cudaStreamCreateWithFlags(&Q.stream[i], cudaStreamNonBlocking));
            CVGS_CUDA(cudaEventCreateWithFlags(&Q.done[i], cudaEventDisableTiming));
        }
        if (Q.start) { cudaEventDestroy(Q.start); Q.start = nullptr; }
        CVGS_CUDA(cudaEventCreateWithFlags(&Q.start, cudaEventDisableTiming));
        Q.device = device;
    }
    CVGS_CUDA(cudaEventRecord(Q.start, stream));
    struct Result {
        int rc = CVGS_OK;
        std::string err;
        int64_t launches = 0;
    };
    Result res[kSeqMaxWorkers];
    auto run = [&](int w, cudaStream_t st) {
        t_items_hint = 8;
        t_in_frame_loop = true;  // the loop owns its streams between fork and join: consecutive launches may overlap
        int rc = CVGS_OK;
        for (int i = 0; i < steps && rc == CVGS_OK; ++i) {
            const int s = i % n_sets;
            if (s % workers != w) continue;
            rc = cvgs_b200_preproc_launch_ex(crops[s], parents[s], n_planes[s], used[s], pipelines[s], st);
        }
        t_items_hint = 0;
        t_in_frame_loop = false;
        return rc;
    };
    seq_pool()->run(
        workers,
        [&](int w) {
            Result& r = res[w];
            if (cudaSetDevice(device) != cudaSuccess || cudaStreamWaitEvent(Q.stream[w], Q.start, 0) != cudaSuccess) {
                r.rc = CVGS_ERR_INVALID_VALUE;
                r.err = "sequence helper: cannot attach to the device";
                return;
            }
            const int64_t before = t_launch_count;
            r.rc = run(w, Q.stream[w]);
            if (r.rc != CVGS_OK) r.err = t_last_error;
            r.launches = t_launch_count - before;
            if (cudaEventRecord(Q.done[w], Q.stream[w]) != cudaSuccess && r.rc == CVGS_OK) {
                r.rc = CVGS_ERR_INVALID_VALUE;
                r.err = "sequence helper: cudaEventRecord failed";
            }
        },
        [&] {
            res[0].rc = run(0, stream);
            if (res[0].rc != CVGS_OK) res[0].err = t_last_error;
        });
    int rc = res[0].rc;
    for (int w = 1; w < workers; ++w) {
        // join the helper stream even after an error: whatever it launched must finish before the caller's stream goes on
        const cudaError_t e = cudaStreamWaitEvent(stream, Q.done[w], 0);
        t_launch_count += res[w].launches;
        if (rc == CVGS_OK && res[w].rc != CVGS_OK) {
            rc = res[w].rc;
            res[0].err = res[w].err;
        }
        if (rc == CVGS_OK && e != cudaSuccess) {
            rc = static_cast<int>(e);
            res[0].err = std::string("cudaStreamWaitEvent: ") + cudaGetErrorString(e);
        }
    }
    if (rc != CVGS_OK) t_last_error = res[0].err;
    return rc;
}

int cvgs_b200_preproc_host_sequence(const void* const* host_images, int32_t image_width, int32_t image_height,
                                    int32_t image_pitch, const cvgs_rect_t* const* rects, const int32_t* n_planes,
                                    const int32_t* used, const cvgs_pipeline_t* const* pipelines,
                                    float* const* host_outs, int32_t n_sets, int32_t steps, void* stream_) {
    CVGS_RANGE("cvgs_b200_preproc_host_sequence");
    if (!host_images || !rects || !n_planes || !used || !pipelines || !host_outs || n_sets <= 0 || steps < 0)
        return fail(CVGS_ERR_INVALID_VALUE, "bad sequence arguments");
    // The frames are independent (each has its own host image and host tensor), so the loop keeps kHostLanes of
    // them in flight on internal streams: the upload of frame i+1 overlaps the kernel of frame i and the download
    // of frame i-1 (two copy engines + SMs).  Seen from the caller's stream the loop is one operation: it starts
    // after everything queued before it and everything queued after it waits for its last frame.
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    int device = 0;
    CVGS_CUDA(cudaGetDevice(&device));
    HostLanes& L = t_ctx.lanes;
    if (L.device != device) {
        for (int i = 0; i < kHostLanes; ++i) {
            if (L.stream[i]) { cudaStreamDestroy(L.stream[i]); L.stream[i] = nullptr; }
            if (L.done[i]) { cudaEventDestroy(L.done[i]); L.done[i] = nullptr; }
            CVGS_CUDA(cudaStreamCreateWithFlags(&L.stream[i], cudaStreamNonBlocking));
            CVGS_CUDA(cudaEventCreateWithFlags(&L.done[i], cudaEventDisableTiming));
        }
        if (L.start) { cudaEventDestroy(L.start); L.start = nullptr; }
        CVGS_CUDA(cudaEventCreateWithFlags(&L.start, cudaEventDisableTiming));
        L.device = device;
    }
    CVGS_CUDA(cudaEventRecord(L.start, stream));
    for (int i = 0; i < kHostLanes; ++i) CVGS_CUDA(cudaStreamWaitEvent(L.stream[i], L.start, 0));
    int rc = CVGS_OK;
    for (int i = 0; i < steps && rc == CVGS_OK; ++i) {
        const int s = i % n_sets, lane = i % kHostLanes;
        rc = preproc_host_impl(t_ctx.lane_buf[lane], host_images[s], image_width, image_height, image_pitch, rects[s],
                               n_planes[s], used[s], pipelines[s], host_outs[s], L.stream[lane],
                               i > 0 ? L.done[(i - 1) % kHostLanes] : nullptr, L.done[lane]);
    }
    // every lane ends with a recorded download (or was never used): the caller's stream resumes after all of them
    for (int i = 0; i < kHostLanes && i < steps; ++i) CVGS_CUDA(cudaStreamWaitEvent(stream, L.done[i], 0));
    return rc;
}

}  // extern "C"
