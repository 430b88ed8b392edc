// oracle/chain_kernels.cu -- TEST / BENCHMARK INFRASTRUCTURE, NOT PRODUCT CODE.
//
// The "OpenCV-CUDA-equivalent" multi-kernel chain of BASELINE.md (baseline M): what the reference's README times its
// fused kernel against (reference README.md:91-97) -- per crop
//     cv::cuda::resize(8UC3 -> 8UC3, INTER_LINEAR)  ->  convertTo(CV_32FC3, alpha)  ->  [cvtColor RGB2BGR]
//     ->  subtract(Scalar)  ->  divide(Scalar)  ->  split
// i.e. 5 (6 with the colour swap) kernel launches per crop, 250-300 per 50-crop frame, every intermediate image
// written to and read back from memory.  Real OpenCV-CUDA cannot be installed here (no sources, no network): these
// kernels RESTATE its arithmetic with the semantics the reference's path has (source
// coordinate = dst * scale without half-pixel centre, taps clamped at the right/bottom edge, FMUL + 3 FFMA, the
// 8-bit resize result rounded to nearest-even and saturated, every later operation rounded on its own, IEEE
// division).  Its output therefore equals the product in (fp_contract = SEPARATE, interp_mode = ROUND_U8) bit for
// bit -- tests/test_chain_gpu.py holds both to that -- and it is timed by bench.py as a labelled baseline.
#include <cuda_runtime.h>
#include <cstdint>

namespace {

__global__ void k_resize_8uc3(const uint8_t* src, int pitch, int sw, int sh, uint8_t* dst, int W, int H, float fx, float fy) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= W || y >= H) return;
    const float sx = __fmul_rn((float)x, fx), sy = __fmul_rn((float)y, fy);
    const int x1 = __float2int_rd(sx), y1 = __float2int_rd(sy);
    const int x2 = x1 + 1, y2 = y1 + 1;
    const int x2r = min(x2, sw - 1), y2r = min(y2, sh - 1);
    const float wx0 = __fsub_rn((float)x2, sx), wx1 = __fsub_rn(sx, (float)x1);
    const float wy0 = __fsub_rn((float)y2, sy), wy1 = __fsub_rn(sy, (float)y1);
    const float w00 = __fmul_rn(wx0, wy0), w10 = __fmul_rn(wx1, wy0), w01 = __fmul_rn(wx0, wy1), w11 = __fmul_rn(wx1, wy1);
    const uint8_t* r0 = src + (size_t)y1 * pitch;
    const uint8_t* r1 = src + (size_t)y2r * pitch;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        float t = __fmul_rn((float)r0[3 * x2r + c], w10);
        t = __fmaf_rn((float)r0[3 * x1 + c], w00, t);
        t = __fmaf_rn((float)r1[3 * x1 + c], w01, t);
        t = __fmaf_rn((float)r1[3 * x2r + c], w11, t);
        const unsigned u = __float2uint_rn(t);
        dst[((size_t)y * W + x) * 3 + c] = (uint8_t)(u > 255u ? 255u : u);
    }
}
__global__ void k_convert_scale(const uint8_t* src, float* dst, int n, float a0, float a1, float a2) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int c = i % 3;
    dst[i] = __fmul_rn((float)src[i], c == 0 ? a0 : c == 1 ? a1 : a2);
}
__global__ void k_swap_rb(const float* src, float* dst, int npix) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= npix) return;
    dst[3 * i] = src[3 * i + 2];
    dst[3 * i + 1] = src[3 * i + 1];
    dst[3 * i + 2] = src[3 * i];
}
__global__ void k_sub_scalar(const float* src, float* dst, int n, float s0, float s1, float s2) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int c = i % 3;
    dst[i] = __fsub_rn(src[i], c == 0 ? s0 : c == 1 ? s1 : s2);
}
__global__ void k_div_scalar(const float* src, float* dst, int n, float d0, float d1, float d2) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int c = i % 3;
    dst[i] = __fdiv_rn(src[i], c == 0 ? d0 : c == 1 ? d1 : d2);
}
__global__ void k_split3(const float* src, float* p0, float* p1, float* p2, int npix) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= npix) return;
    p0[i] = src[3 * i];
    p1[i] = src[3 * i + 1];
    p2[i] = src[3 * i + 2];
}

}  // namespace

extern "C" {

// Bytes of scratch memory one call needs (per crop buffers are reused from crop to crop, as the README loop does).
size_t chain_workspace_bytes(int W, int H) { return (size_t)W * H * 3 * (1 + 4 + 4); }

// n crops -> out[n][3][H][W]; the mul/sub/div constants are given in the channel order of the tensor they apply to
// (after the optional swap), like the cv::Scalar arguments of the README chain.  Returns the number of launches.
int chain_preproc(const void* const* ptrs, const int* ws, const int* hs, const int* pitches, int n, int W, int H, int swap_rb,
                  const float* mul, const float* sub, const float* div, float* out, void* workspace, void* stream_) {
    cudaStream_t st = (cudaStream_t)stream_;
    uint8_t* u8 = (uint8_t*)workspace;
    float* fa = (float*)(u8 + (((size_t)W * H * 3 + 255) / 256 * 256));
    float* fb = fa + (size_t)W * H * 3;
    const int npix = W * H, nel = npix * 3;
    const dim3 b2(32, 8), g2((W + 31) / 32, (H + 7) / 8);
    const int b1 = 256, g1e = (nel + 255) / 256, g1p = (npix + 255) / 256;
    int launches = 0;
    for (int i = 0; i < n; ++i) {
        const float fx = (float)(1.0 / ((double)W / (double)ws[i])), fy = (float)(1.0 / ((double)H / (double)hs[i]));
        k_resize_8uc3<<<g2, b2, 0, st>>>((const uint8_t*)ptrs[i], pitches[i], ws[i], hs[i], u8, W, H, fx, fy);
        // convertTo(alpha) happens before the swap in the chain: its constants are in source order
        const float a0 = swap_rb ? mul[2] : mul[0], a2 = swap_rb ? mul[0] : mul[2];
        k_convert_scale<<<g1e, b1, 0, st>>>(u8, fa, nel, a0, mul[1], a2);
        float* cur = fa;
        float* other = fb;
        launches += 2;
        if (swap_rb) {
            k_swap_rb<<<g1p, b1, 0, st>>>(cur, other, npix);
            float* t = cur; cur = other; other = t;
            ++launches;
        }
        k_sub_scalar<<<g1e, b1, 0, st>>>(cur, other, nel, sub[0], sub[1], sub[2]);
        k_div_scalar<<<g1e, b1, 0, st>>>(other, cur, nel, div[0], div[1], div[2]);
        float* plane = out + (size_t)i * 3 * npix;
        k_split3<<<g1p, b1, 0, st>>>(cur, plane, plane + npix, plane + 2 * (size_t)npix, npix);
        launches += 3;
    }
    return cudaGetLastError() == cudaSuccess ? launches : -1;
}

// Frame loop in native code: `steps` consecutive chain_preproc calls, call i using argument set i % n_sets.
int chain_preproc_sequence(const void* const* const* ptrs, const int* const* ws, const int* const* hs, const int* const* pitches,
                           int n, int W, int H, int swap_rb, const float* mul, const float* sub, const float* div,
                           float* const* outs, void* workspace, int n_sets, int steps, void* stream) {
    long long launches = 0;
    for (int i = 0; i < steps; ++i) {
        const int s = i % n_sets;
        const int rc = chain_preproc(ptrs[s], ws[s], hs[s], pitches[s], n, W, H, swap_rb, mul, sub, div, outs[s], workspace, stream);
        if (rc < 0) return -1;
        launches += rc;
    }
    return (int)(launches > 0x7fffffff ? 0x7fffffff : launches);
}

}  // extern "C"
