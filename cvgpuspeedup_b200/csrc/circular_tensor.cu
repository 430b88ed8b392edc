// circular_tensor.cu -- the CircularTensor shift + process path as ONE sm_100a kernel.
//
// Replaces fk::CircularTensor::update (reference fkl/include/fused_kernel/core/data/circular_tensor.cuh:111-146),
// which launches launchDivergentBatchTransformDPP_Kernel with blockIdx.z selecting, per plane, either
// the user's pipeline (new frame) or a scalar 4-byte plane copy (data_parallel_patterns.cuh:221-254).
//
// Same contract: after every update data() is a dense, time-ordered tensor at a stable address.
// State = the public tensor + a ring of the same shape + the next ring slot
// (the reference's m_tempTensor / m_nextUpdateIdx, circular_tensor.cuh:148-150).
//
// Kernel roles (one launch, role chosen per CTA):
//   compute CTAs  run resize + op chain on the new frame and write the plane twice
//                 (ring slot + newest position of the public tensor);
//   copy CTAs     move the other BATCH-1 planes ring -> public with the rotated index, 16 bytes per
//                 thread access, 4 independent accesses in flight per thread.
#include <cuda_runtime.h>

#include <new>

#include "cvgs_device.cuh"
#include "cvgs_runtime.hpp"
#include "preproc_direct.cuh"
#include "preproc_host.hpp"
#include "preproc_tma.cuh"  // specialize_division

namespace cvgs {

struct CtParams {
    PreprocParams pre;       // pre.out describes the PUBLIC tensor; z of the new plane = upd
    float* ring;             // ring tensor, same strides
    const float* ring_ro;
    int32_t batch, order;    // CVGS_CT_NEWEST_FIRST / OLDEST_FIRST
    int32_t first;           // m_nextUpdateIdx before this update
    int32_t upd;             // plane of the public tensor that receives the new frame
    int32_t slot;            // ring slot that receives the new frame
    int32_t compute_ctas;    // CTAs [0, compute_ctas) compute, the rest copy
    int32_t bw_log2;
    int32_t tiles_x;
    long long plane_px;      // W*H
    int32_t copy_vec4;       // copy units are 16-byte aligned multiples of 4 floats
    int32_t copy_chunks;     // copy CTAs per copy unit when copy_vec4
    int32_t units_per_plane; // copy units per batch plane: the colour planes (split tensors) or 1 (packed elements)
    long long unit_floats;   // floats per copy unit: W*H (split) or channels * W*H (packed)
};

constexpr int kCtChunk = 8192;  // float4 per copy CTA (128 KB): 256 threads x 4 accesses x 8 rounds

__device__ __forceinline__ int ring_source(const CtParams& K, int z) {
    // computeCircularThreadIdx, reference memory_operations.cuh:388-399
    if (K.order == CVGS_CT_NEWEST_FIRST) {
        const int s = K.first - z;
        return s < 0 ? s + K.batch : s;
    }
    const int s = z + K.first;
    return s >= K.batch ? s - K.batch : s;
}

// NC = channels of the source pixel, NR = registers per pixel of the chain (4 when a 3-channel source gains an alpha).
template <int NC, int NR>
__global__ void __launch_bounds__(256) circular_update_kernel(const __grid_constant__ CtParams K,
                                                              const __grid_constant__ DevCrop frame) {
    const PreprocParams& P = K.pre;
    if ((int)blockIdx.x < K.compute_ctas) {
        // ---- compute role: one tile of the new plane ----
        const int tid = threadIdx.x;
        const int tx = tid & ((1 << K.bw_log2) - 1);
        const int ty = tid >> K.bw_log2;
        const int bx = blockIdx.x % K.tiles_x, by = blockIdx.x / K.tiles_x;
        const int x0 = ((bx << K.bw_log2) + tx) * 4;
        const int y = by * (256 >> K.bw_log2) + ty;
        if (x0 >= P.W || y >= P.H) return;
        const int nvalid = min(4, P.W - x0);
        float v[4][NC];
        gather_quad(P, frame, y, x0, nvalid, v);
        float r[4][NR];
#pragma unroll
        for (int p = 0; p < 4; ++p)
#pragma unroll
            for (int c = 0; c < NR; ++c) r[p][c] = c < NC ? v[p][c] : 0.f;
        apply_program<4, NR>(P.prog, r);
        store_pixels<4, NR>(P, K.upd, y, x0, nvalid, r);  // public tensor, newest position
        PreprocParams R = P;                              // same strides, ring base
        R.out.base = K.ring;
        store_pixels<4, NR>(R, K.slot, y, x0, nvalid, r);
        return;
    }
    // ---- copy role: planes z != upd, public[z][c] <- ring[ring_source(z)][c] ----
    const int ncopy = gridDim.x - K.compute_ctas;
    const int cta = blockIdx.x - K.compute_ctas;
    const OutDesc& o = P.out;
    const int upp = K.units_per_plane;
    const int units = (K.batch - 1) * upp;  // (plane, colour) pairs, or whole planes of packed elements
    if (K.copy_vec4) {
        // CTA -> (unit, chunk): one division per CTA, then 16-byte accesses at a fixed stride, 4 in flight per thread
        const int n4 = (int)(K.unit_floats / 4);
        const int u = cta / K.copy_chunks;
        const int chunk = cta - u * K.copy_chunks;
        if (u >= units) return;
        int z = u / upp;
        const int c = u - z * upp;
        if (z >= K.upd) ++z;  // skip the plane the compute CTAs write
        const int zs = ring_source(K, z);
        const float4* src = reinterpret_cast<const float4*>(K.ring_ro + zs * o.z_stride + c * o.c_stride);
        float4* dst = reinterpret_cast<float4*>(o.base + z * o.z_stride + c * o.c_stride);
        const int lo = chunk * kCtChunk, hi = min(n4, lo + kCtChunk);
        for (int i = lo + threadIdx.x; i < hi; i += 4 * 256) {
            float4 r[4];
#pragma unroll
            for (int k = 0; k < 4; ++k)
                if (i + k * 256 < hi) r[k] = __ldcs(src + i + k * 256);
#pragma unroll
            for (int k = 0; k < 4; ++k)
                if (i + k * 256 < hi) __stcs(dst + i + k * 256, r[k]);
        }
    } else {
        const long long total = K.unit_floats * units;
        const long long stride = (long long)ncopy * 256;
        for (long long j = (long long)cta * 256 + threadIdx.x; j < total; j += stride) {
            const int u = (int)(j / K.unit_floats);
            const long long e = j - (long long)u * K.unit_floats;
            int z = u / upp;
            const int c = u - z * upp;
            if (z >= K.upd) ++z;
            const int zs = ring_source(K, z);
            o.base[z * o.z_stride + c * o.c_stride + e] = K.ring_ro[zs * o.z_stride + c * o.c_stride + e];
        }
    }
}

struct CircularTensor {
    int32_t w, h, cp, ec, batch, order, mode, device;  // cp colour planes of ec-channel elements
    float* pub = nullptr;
    float* ring = nullptr;
    int32_t next = 0;
};

}  // namespace cvgs

using namespace cvgs;

extern "C" {

int cvgs_b200_ct_create_ex(void** handle, int32_t width, int32_t height, int32_t color_planes, int32_t elem_channels,
                           int32_t batch, int32_t order, int32_t plane_mode, int32_t device) {
    if (!handle) return fail(CVGS_ERR_INVALID_VALUE, "handle is NULL");
    *handle = nullptr;
    if (width <= 0 || height <= 0 || batch <= 0) return fail(CVGS_ERR_INVALID_VALUE, "bad tensor shape");
    // split tensors: 1, 3 or 4 colour planes of floats; packed tensors: one plane of float3 / float4 elements
    const bool split = elem_channels == 1 && (color_planes == 1 || color_planes == 3 || color_planes == 4);
    const bool packed = color_planes == 1 && (elem_channels == 3 || elem_channels == 4);
    if (!split && !packed)
        return fail(CVGS_ERR_NOT_SUPPORTED, "CircularTensor: 1, 3 or 4 colour planes of floats, or one plane of 3- / 4-channel float pixels");
    if (order < 0 || order > 1 || plane_mode < 0 || plane_mode > 1) return fail(CVGS_ERR_INVALID_VALUE, "bad order/mode");
    int prev = 0;
    CVGS_CUDA(cudaGetDevice(&prev));
    CVGS_CUDA(cudaSetDevice(device));
    CircularTensor* t = new (std::nothrow) CircularTensor();
    if (!t) return fail(2 /*cudaErrorMemoryAllocation*/, "out of host memory");
    t->w = width; t->h = height; t->cp = color_planes; t->ec = elem_channels; t->batch = batch; t->order = order; t->mode = plane_mode;
    t->device = device;
    const size_t bytes = sizeof(float) * (size_t)width * height * color_planes * elem_channels * batch;
    cudaError_t e = cudaMalloc(reinterpret_cast<void**>(&t->pub), bytes);
    if (e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void**>(&t->ring), bytes);
    if (e == cudaSuccess) e = cudaMemset(t->pub, 0, bytes);
    if (e == cudaSuccess) e = cudaMemset(t->ring, 0, bytes);
    cudaSetDevice(prev);
    if (e != cudaSuccess) {
        cudaFree(t->pub); cudaFree(t->ring);
        delete t;
        return cuda_fail(e, "CircularTensor allocation");
    }
    *handle = t;
    return CVGS_OK;
}

int cvgs_b200_ct_create(void** handle, int32_t width, int32_t height, int32_t color_planes, int32_t batch,
                        int32_t order, int32_t plane_mode, int32_t device) {
    return cvgs_b200_ct_create_ex(handle, width, height, color_planes, 1, batch, order, plane_mode, device);
}

int cvgs_b200_ct_update(void* handle, const cvgs_crop_t* frame, const cvgs_pipeline_t* pipeline, void* stream_) {
    CVGS_RANGE("cvgs_b200_ct_update");
    CircularTensor* t = static_cast<CircularTensor*>(handle);
    if (!t) return fail(CVGS_ERR_INVALID_VALUE, "handle is NULL");
    if (!frame) return fail(CVGS_ERR_INVALID_VALUE, "frame is NULL");
    int cur_device = -1;
    CVGS_CUDA(cudaGetDevice(&cur_device));
    if (cur_device != t->device)
        return fail(CVGS_ERR_INVALID_VALUE, "CircularTensor lives on device " + std::to_string(t->device) + ", the current device is " +
                                                std::to_string(cur_device));
    if (int rc = validate_pipeline(pipeline)) return rc;
    if (pipeline->dst_type == CVGS_8UC3 || pipeline->dst_type == CVGS_8UC4 || pipeline->out_row_pitch != 0) return fail(CVGS_ERR_NOT_SUPPORTED, "CircularTensor planes are float");
    if (CVGS_IS_YUV(pipeline->src_type)) return fail(CVGS_ERR_NOT_SUPPORTED, "CircularTensor takes CV_8U / CV_16U / CV_16S frames");
    if (pipeline->dst_width != t->w || pipeline->dst_height != t->h)
        return fail(CVGS_ERR_INVALID_VALUE, "pipeline destination size must equal the tensor plane size");
    cvgs_pipeline_t p = *pipeline;
    // packed elements: [z][y][x][channel] (TensorWrite); split: one plane per channel, plane-major or colour-major
    p.out_layout = t->ec > 1 ? CVGS_OUT_NHWC : (t->mode == CVGS_CT_STANDARD ? CVGS_OUT_NCHW : CVGS_OUT_CNHW);
    p.out_plane_stride = 0;
    p.dst_type = 0;
    CtParams K;
    std::memset(&K, 0, sizeof K);
    if (int rc = build_params(p, t->batch, 1, t->pub, K.pre)) return rc;
    if (K.pre.prog.nc_out != t->cp * t->ec)
        return fail(CVGS_ERR_INVALID_VALUE, "the chain produces " + std::to_string(K.pre.prog.nc_out) + " channels, the tensor stores " +
                                                std::to_string(t->cp * t->ec));
    specialize_division(K.pre.prog, K.pre.nc, K.pre.bg);
    DevCrop dc;
    if (int rc = fill_crop(*frame, p, 0, dc)) return rc;
    K.ring = t->ring;
    K.ring_ro = t->ring;
    K.batch = t->batch;
    K.order = t->order;
    K.first = t->next;
    K.upd = t->order == CVGS_CT_NEWEST_FIRST ? 0 : t->batch - 1;
    K.slot = (K.upd + t->next) % t->batch;
    K.plane_px = (long long)t->w * t->h;
    K.units_per_plane = t->ec > 1 ? 1 : t->cp;
    K.unit_floats = K.plane_px * (t->ec > 1 ? t->ec : 1);
    K.copy_vec4 = (K.unit_floats % 4) == 0;  // cudaMalloc bases are 256-byte aligned

    // compute tiles: same block shape heuristic as the batch kernel
    const int qw = (t->w + 3) / 4;
    int l = 3, best_pad = 1 << 30;
    for (int c = 3; c <= 6; ++c) {
        const int bw = 1 << c, padded = (qw + bw - 1) / bw * bw;
        if (padded <= best_pad) { best_pad = padded; l = c; }
    }
    const int rows = 256 >> l;
    K.bw_log2 = l;
    K.tiles_x = (qw + (1 << l) - 1) >> l;
    const long long tiles_y = (t->h + rows - 1) / rows;
    const long long compute = (long long)K.tiles_x * tiles_y;
    if (compute > (1ll << 30)) return fail(CVGS_ERR_INVALID_VALUE, "plane too large");
    K.compute_ctas = (int)compute;
    int copy_ctas = 0;
    if (t->batch > 1 && K.copy_vec4) {
        K.copy_chunks = (int)((K.unit_floats / 4 + kCtChunk - 1) / kCtChunk);
        copy_ctas = K.copy_chunks * K.units_per_plane * (t->batch - 1);
    } else if (t->batch > 1) {
        const long long work = K.unit_floats * K.units_per_plane * (t->batch - 1);  // thread-iterations
        const long long want = (work + 255) / 256;
        const long long cap = (long long)sm_count_of(t->device) * 8;
        copy_ctas = (int)std::max<long long>(1, std::min(want, cap));
    }
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    const int grid = K.compute_ctas + copy_ctas;
    const int nc = K.pre.nc, nr = K.pre.prog.nregs;
    if (nc == 3 && nr == 4) circular_update_kernel<3, 4><<<grid, 256, 0, stream>>>(K, dc);
    else if (nc == 4) circular_update_kernel<4, 4><<<grid, 256, 0, stream>>>(K, dc);
    else circular_update_kernel<3, 3><<<grid, 256, 0, stream>>>(K, dc);
    count_launch();
    CVGS_CUDA(cudaGetLastError());
    t->next = (t->next + 1) % t->batch;  // circular_tensor.cuh:144
    return CVGS_OK;
}

void* cvgs_b200_ct_data(void* handle) {
    CircularTensor* t = static_cast<CircularTensor*>(handle);
    return t ? t->pub : nullptr;
}

int cvgs_b200_ct_destroy(void* handle) {
    CircularTensor* t = static_cast<CircularTensor*>(handle);
    if (!t) return CVGS_OK;
    cudaError_t e1 = cudaFree(t->pub), e2 = cudaFree(t->ring);
    delete t;
    if (e1 != cudaSuccess) return cuda_fail(e1, "cudaFree(public tensor)");
    if (e2 != cudaSuccess) return cuda_fail(e2, "cudaFree(ring tensor)");
    return CVGS_OK;
}

}  // extern "C"
