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

// -------------------------------------------------------------------------------------------------------------------
// Fast instantiation: float tensor output (planar rows), chain = fma(v, a, b) [then the two-operation division by the
// launch constants, specialize_division], i.e. every chain the reference's warp tests and the detector pipelines run.
//   * a warp owns 128 x-adjacent pixels of one row as four groups of 32: lane l takes pixels l, l + 32, l + 64, l + 96, so
//     one load instruction of the warp taps 32 ADJACENT destination pixels -- a few sectors of one or two source rows --
//     where the four-pixels-per-lane form above spreads it over four times the span (10.8 sectors per request measured,
//     r01f_warp_kernel_ncu_full.txt); stores are 128-byte warp rows.
//   * no interpreter: the chain's constants sit in the kernel parameters, the descriptor is read once per thread.
//   * no branch around the taps: a pixel outside the source image reads the image's first pixel and is replaced by 0.
// Same rounding sequence as warp_one_pixel, so both kernels give the same bits (tests/test_warp_gpu.py runs both).
struct WarpFastParams {
    int32_t W, H, used;
    int32_t rows;        // rows per warp (y, y + 8, ...): a CTA covers 8 * rows rows
    float bg[4];
    float a[4], b[4];    // v = fma(v, a, b), indexed by source channel
    float zh[4], zl[4];  // DIV: v = fma(v, zh, v * zl)
    float* base;
    long long z_stride, row_stride;  // floats
    long long c_off[4];              // source channel r goes to row + c_off[r]
};

// sample -> float on the ALU pipe (I2FP): left to itself the compiler converts an 8- / 16-bit load with I2F.U16 / .S16 on
// the quarter-rate conversion pipe, twelve per pixel.  The load is predicated: a pixel outside the source image loads
// nothing (its value is replaced by 0 afterwards), so its tap addresses need no clamping.
template <typename T>
__device__ __forceinline__ float warp_sample(const T* p, int pred) {
    float f;
    if (sizeof(T) == 1) {
        unsigned u;
        asm volatile("{\n.reg .pred q;\nsetp.ne.s32 q, %2, 0;\n@q ld.global.nc.u8 %0, [%1];\n}" : "=r"(u) : "l"(p), "r"(pred));
        asm("cvt.rn.f32.u32 %0, %1;" : "=f"(f) : "r"(u));
    } else if (T(-1) < T(0)) {
        int i;
        asm volatile("{\n.reg .pred q;\nsetp.ne.s32 q, %2, 0;\n@q ld.global.nc.s16 %0, [%1];\n}" : "=r"(i) : "l"(p), "r"(pred));
        asm("cvt.rn.f32.s32 %0, %1;" : "=f"(f) : "r"(i));
    } else {
        unsigned u;
        asm volatile("{\n.reg .pred q;\nsetp.ne.s32 q, %2, 0;\n@q ld.global.nc.u16 %0, [%1];\n}" : "=r"(u) : "l"(p), "r"(pred));
        asm("cvt.rn.f32.u32 %0, %1;" : "=f"(f) : "r"(u));
    }
    return f;
}

// DevWarp::pad of the fast kernel: how the source coordinate is formed
enum WarpMode : int32_t {
    WM_AFFINE = 0,
    WM_PERSPECTIVE_NORMAL = 1,  // the host has proven every denominator of the plane finite with magnitude in [2^-100, 2^100]:
                                // the correctly rounded reciprocal is MUFU.RCP + one Newton step, no range check
    WM_PERSPECTIVE = 2,
};

#ifndef CVGS_WARP_MINB
#define CVGS_WARP_MINB 5
#endif
#ifndef CVGS_WARP_UNROLL
#define CVGS_WARP_UNROLL 1
#endif
// Rows y, y + 8, ... (K.rows rows per warp) x four pixels per row: columns xb, xb + 32, xb + 64, xb + 96 (xb = first column of
// the warp's span + lane), each through the chain and into its planes.  The four pixels are unrolled -- store offsets are
// immediates, the chain's constants and the matrix stay in registers -- but kept in program order by a dependency of each
// pixel's x on the previous pixel's stores (the empty asm): left free, the compiler hoists all 48 loads and the kernel
// runs at 80-100 registers (two or three CTAs per SM); the kernel is issue-bound, so occupancy pays and ILP does not.
template <typename T, int NC, bool DIV, int MODE>
__device__ __forceinline__ void warp_fast_pixels(const WarpFastParams& K, const DevWarp& d, int xb, int y, const float* plane0) {
    const float m0 = d.m[0], m1 = d.m[1], m2 = d.m[2], m3 = d.m[3], m4 = d.m[4], m5 = d.m[5], m6 = d.m[6], m7 = d.m[7], m8 = d.m[8];
    const unsigned w = static_cast<unsigned>(d.w), h = static_cast<unsigned>(d.h);
    const int wm1 = d.w - 1, hm1 = d.h - 1;
    int pitch = d.pitch;
    const uint8_t* base = d.data;
    constexpr int kPx = NC * sizeof(T);
    const int np = min(4, (K.W - xb + 31) >> 5);
    const float fx0 = static_cast<float>(xb);
    // values the pixel loop needs in vector registers (the row pitch as an IMAD.WIDE factor, the chain's addends as FFMA
    // operands next to a constant-bank factor): passed through an empty asm so that they stay there instead of being
    // re-read from the constant bank for every pixel
    float cb[NC];
#pragma unroll
    for (int c = 0; c < NC; ++c) {
        cb[c] = K.b[c];
        asm volatile("" : "+f"(cb[c]));
    }
    asm volatile("" : "+r"(pitch));
    int four = 4;
    asm volatile("" : "+r"(four));
    // plane c of this batch plane starts at pl[c] (uniform over the CTA); a store address is pl[c] + 4 * (offset in the plane)
    const float* pl[NC];
#pragma unroll
    for (int c = 0; c < NC; ++c) pl[c] = plane0 + K.c_off[c];
#pragma unroll 1
    for (int r = 0; r < K.rows && y < K.H; ++r, y += 8) {
        const float fy = static_cast<float>(y);
        const float tx = __fmul_rn(m1, fy), ty = __fmul_rn(m4, fy), tz = __fmul_rn(m7, fy);
        float fx = fx0;
        int idx = y * static_cast<int>(K.row_stride) + xb;  // float offset inside a plane (planes are below 2^31 bytes: host check)
        constexpr int kUnroll = CVGS_WARP_UNROLL;
#pragma unroll kUnroll
        for (int p = 0; p < np; ++p) {
            {
                float sx = __fadd_rn(__fmaf_rn(m0, fx, tx), m2);
                float sy = __fadd_rn(__fmaf_rn(m3, fx, ty), m5);
                if (MODE != WM_AFFINE) {
                    const float den = __fadd_rn(__fmaf_rn(m6, fx, tz), m8);
                    float coeff;
                    if (MODE == WM_PERSPECTIVE_NORMAL) {  // what __frcp_rn executes for an operand in the normal range
                        float rr;
                        asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rr) : "f"(den));
                        const float e = __fmaf_rn(den, rr, -1.0f);
                        coeff = __fmaf_rn(rr, -e, rr);
                    } else {
                        coeff = __frcp_rn(den);
                    }
                    sx = __fmul_rn(coeff, sx);
                    sy = __fmul_rn(coeff, sy);
                }
                // 0 <= sx < w  <=>  floor(sx) in [0, w) as an unsigned compare (the conversion saturates; -0 floors to 0 like
                // the reference's sx >= 0; w, h <= 2^24 so float(w) is exact) -- except NaN, which converts to 0: tested apart
                const int x1 = __float2int_rd(sx), y1 = __float2int_rd(sy);
                int num;
                asm("{\n.reg .pred q;\nsetp.num.f32 q, %1, %2;\nselp.s32 %0, 1, 0, q;\n}" : "=r"(num) : "f"(sx), "f"(sy));
                const int in = static_cast<unsigned>(x1) < w && static_cast<unsigned>(y1) < h && num;
                const float fx1 = static_cast<float>(x1), fy1 = static_cast<float>(y1);
                const float wx1 = __fsub_rn(sx, fx1), wx0 = __fsub_rn(__fadd_rn(fx1, 1.0f), sx);  // float(x1 + 1) == fx1 + 1 below 2^24
                const float wy1 = __fsub_rn(sy, fy1), wy0 = __fsub_rn(__fadd_rn(fy1, 1.0f), sy);
                const float w00 = __fmul_rn(wx0, wy0), w10 = __fmul_rn(wx1, wy0);
                const float w01 = __fmul_rn(wx0, wy1), w11 = __fmul_rn(wx1, wy1);
                const int x2r = min(x1 + 1, wm1), y2r = min(y1 + 1, hm1);
                // the two row pointers are materialised (empty asm): each tap pointer is then one IMAD.WIDE on top of them
                // instead of a 64-bit offset plus an IADD3 / IADD3.X pair that adds the (uniform) image base
                const uint8_t* r0 = base + static_cast<long long>(y1) * pitch;
                const uint8_t* r1 = base + static_cast<long long>(y2r) * pitch;
                asm volatile("" : "+l"(r0));
                asm volatile("" : "+l"(r1));
                const T* p00 = reinterpret_cast<const T*>(r0 + static_cast<long long>(x1) * kPx);
                const T* p10 = reinterpret_cast<const T*>(r0 + static_cast<long long>(x2r) * kPx);
                const T* p01 = reinterpret_cast<const T*>(r1 + static_cast<long long>(x1) * kPx);
                const T* p11 = reinterpret_cast<const T*>(r1 + static_cast<long long>(x2r) * kPx);
#pragma unroll
                for (int c = 0; c < NC; ++c) {
                    float t = __fmul_rn(warp_sample(p10 + c, in), w10);
                    t = __fmaf_rn(warp_sample(p00 + c, in), w00, t);
                    t = __fmaf_rn(warp_sample(p01 + c, in), w01, t);
                    t = __fmaf_rn(warp_sample(p11 + c, in), w11, t);
                    t = in ? t : 0.f;
                    t = __fmaf_rn(t, K.a[c], cb[c]);
                    if (DIV) t = __fmaf_rn(t, K.zh[c], __fmul_rn(t, K.zl[c]));
                    // one IMAD.WIDE (uniform plane base + 4 * idx) per store instead of a 64-bit pointer kept per plane
                    // (the factor 4 sits in a register: IMAD.WIDE takes one uniform operand, the plane base)
                    asm volatile("{\n.reg .u64 a;\nmad.wide.s32 a, %1, %3, %0;\n" CVGS_ST_F32 " [a], %2;\n}" ::"l"(pl[c]), "r"(idx), "f"(t), "r"(four) : "memory");
                }
                idx += 32;
                fx = __fadd_rn(fx, 32.0f);  // exact
                asm volatile("" : "+f"(fx)::"memory");
            }
        }
    }
}

template <typename T, int NC, bool DIV>
__global__ void __launch_bounds__(256, NC == 4 ? 4 : CVGS_WARP_MINB)  // four channels: 64 registers, no spills
preproc_warp_fast_kernel(const __grid_constant__ WarpFastParams K, const __grid_constant__ WarpTable Tb, int z0) {
    const int xb = blockIdx.x * 128 + (threadIdx.x & 31);
    const int y = blockIdx.y * (8 * K.rows) + (threadIdx.x >> 5);
    if (y >= K.H || xb >= K.W) return;
    const int zl = blockIdx.z;
    const int z = z0 + zl;
    const float* plane0 = K.base + z * K.z_stride;  // uniform over the CTA
    if (z < K.used) {
        const DevWarp& d = Tb.w[zl];
        const int mode = d.pad;  // uniform over the CTA
        if (mode == WM_AFFINE) warp_fast_pixels<T, NC, DIV, WM_AFFINE>(K, d, xb, y, plane0);
        else if (mode == WM_PERSPECTIVE_NORMAL) warp_fast_pixels<T, NC, DIV, WM_PERSPECTIVE_NORMAL>(K, d, xb, y, plane0);
        else warp_fast_pixels<T, NC, DIV, WM_PERSPECTIVE>(K, d, xb, y, plane0);
        return;
    }
    float* row = K.base + z * K.z_stride + y * K.row_stride + xb;
#pragma unroll
    for (int c = 0; c < NC; ++c) {  // unused plane: the chain of the background value
        float t = __fmaf_rn(K.bg[c], K.a[c], K.b[c]);
        if (DIV) t = __fmaf_rn(t, K.zh[c], __fmul_rn(t, K.zl[c]));
        float* dst = row + K.c_off[c];
        for (int r = 0, yy = y; r < K.rows && yy < K.H; ++r, yy += 8, dst += 8 * K.row_stride)
#pragma unroll
            for (int p = 0; p < 4; ++p)
                if (xb + 32 * p < K.W) st_cs_f32(dst + 32 * p, t);
    }
}

}  // namespace cvgs
