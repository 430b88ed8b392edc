// preproc_warp.cuh -- batched affine / perspective warp in front of the op chain (cvgs_b200_warp_launch).
//
// Replaces fk::Warping<WT, PerThreadRead<_2D, uchar3>> under BatchRead (reference fkl/include/fused_kernel/algorithms/
// image_processing/warping.cuh:43-91; cvGS::warp, include/cvGPUSpeedup.cuh:285-442).  The source coordinate of every
// destination pixel is data dependent, so there is no rectangular footprint to stage with TMA: taps are gathered
// straight from global memory (L1/L2 absorb the reuse between neighbours) and the kernel is bound by the output
// stream.  One thread produces four x-adjacent pixels so the planar stores stay 16 bytes wide.
//
// Rounding sequence (reference SASS, sm_100a, nvcc 12.9 -- see oracle/oracle.c warp_pixel):
//   row r of the matrix:   t = FMUL(m[r][1], y);  t = FFMA(m[r][0], x, t);  t = FADD(t, m[r][2])
//   perspective:           c = RCP.RN(row 2);  sx = FMUL(c, row 0);  sy = FMUL(c, row 1)
//   interpolation:         weights from FADDs, products w10, w00, w01, w11 (FMULs), then per channel
//                          FMUL(p10, w10), FFMA(p00, w00), FFMA(p01, w01), FFMA(p11, w11) as in the resize
#pragma once
#include "cvgs_device.cuh"

namespace cvgs {

struct alignas(16) DevWarp {
    const uint8_t* data;
    int32_t w, h, pitch, type;
    float m[9];
    int32_t pad;
};
static_assert(sizeof(DevWarp) == 64, "DevWarp is one 64-byte descriptor");

constexpr int kWarpParamPlanes = 48;  // descriptors per launch (kernel parameters); larger batches are chunked
struct WarpTable {
    DevWarp w[kWarpParamPlanes];
};

// T = unsigned char / unsigned short / short, NC = 3 / 4: the pixel types cvGS::warp<WT, InputType> is instantiated with
template <typename T, int NC>
__device__ __forceinline__ void warp_one_pixel(const DevWarp& d, int x, int y, float (&out)[NC]) {
    const float fx = static_cast<float>(x), fy = static_cast<float>(y);
    float sx = __fadd_rn(__fmaf_rn(d.m[0], fx, __fmul_rn(d.m[1], fy)), d.m[2]);
    float sy = __fadd_rn(__fmaf_rn(d.m[3], fx, __fmul_rn(d.m[4], fy)), d.m[5]);
    if (d.type == CVGS_WARP_PERSPECTIVE) {
        const float coeff = __frcp_rn(__fadd_rn(__fmaf_rn(d.m[6], fx, __fmul_rn(d.m[7], fy)), d.m[8]));
        sx = __fmul_rn(coeff, sx);
        sy = __fmul_rn(coeff, sy);
    }
    if (!(sx >= 0.f && sx < static_cast<float>(d.w) && sy >= 0.f && sy < static_cast<float>(d.h))) {
#pragma unroll
        for (int c = 0; c < NC; ++c) out[c] = 0.f;
        return;
    }
    const int x1 = __float2int_rd(sx), y1 = __float2int_rd(sy);
    const int x2 = x1 + 1, y2 = y1 + 1;
    const int x2r = min(x2, d.w - 1), y2r = min(y2, d.h - 1);
    const float wx1 = __fsub_rn(sx, static_cast<float>(x1)), wx0 = __fsub_rn(static_cast<float>(x2), sx);
    const float wy1 = __fsub_rn(sy, static_cast<float>(y1)), wy0 = __fsub_rn(static_cast<float>(y2), sy);
    const float w00 = __fmul_rn(wx0, wy0), w10 = __fmul_rn(wx1, wy0);
    const float w01 = __fmul_rn(wx0, wy1), w11 = __fmul_rn(wx1, wy1);
    const T* r0 = reinterpret_cast<const T*>(d.data + static_cast<long long>(y1) * d.pitch);
    const T* r1 = reinterpret_cast<const T*>(d.data + static_cast<long long>(y2r) * d.pitch);
    const T* p00 = r0 + NC * x1;
    const T* p10 = r0 + NC * x2r;
    const T* p01 = r1 + NC * x1;
    const T* p11 = r1 + NC * x2r;
#pragma unroll
    for (int c = 0; c < NC; ++c) {
        float t = __fmul_rn(static_cast<float>(__ldg(p10 + c)), w10);
        t = __fmaf_rn(static_cast<float>(__ldg(p00 + c)), w00, t);
        t = __fmaf_rn(static_cast<float>(__ldg(p01 + c)), w01, t);
        t = __fmaf_rn(static_cast<float>(__ldg(p11 + c)), w11, t);
        out[c] = t;
    }
}

// grid (ceil(W / (32 * PX)), ceil(H / 8), planes of this chunk); block 256 = 32 lanes x 8 rows; a thread produces PX
// x-adjacent pixels (PX = 4: 16-byte planar stores; PX = 1: most threads in flight, the stores of a warp still coalesce)
template <int PX, typename T = unsigned char, int NC = 3>
__global__ void __launch_bounds__(256)
preproc_warp_kernel(const __grid_constant__ PreprocParams P, const __grid_constant__ WarpTable Tb, int z0) {
    const int x0 = (blockIdx.x * 32 + (threadIdx.x & 31)) * PX;
    const int y = blockIdx.y * 8 + (threadIdx.x >> 5);
    const int zl = blockIdx.z;
    const int z = z0 + zl;
    if (x0 >= P.W || y >= P.H) return;
    const int nvalid = min(PX, P.W - x0);
    float v[PX][NC];
    if (z < P.used) {
        const DevWarp& d = Tb.w[zl];
#pragma unroll
        for (int p = 0; p < PX; ++p) {
            if (p < nvalid) {
                warp_one_pixel<T, NC>(d, x0 + p, y, v[p]);
            } else {
#pragma unroll
                for (int c = 0; c < NC; ++c) v[p][c] = 0.f;
            }
        }
    } else {
#pragma unroll
        for (int p = 0; p < PX; ++p)
#pragma unroll
            for (int c = 0; c < NC; ++c) v[p][c] = P.bg[c];
    }
    apply_program<PX, NC>(P.prog, v);
    store_pixels<PX, NC>(P, z, y, x0, nvalid, v);
}

}  // namespace cvgs
