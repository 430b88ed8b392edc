// preproc_direct.cuh -- direct-gather computation of one quad (4 x-adjacent output pixels):
// bilinear taps are fetched straight from global memory with aligned 32-bit loads.
// Used by preproc_direct_kernel and by the CircularTensor update kernel.
#pragma once
#include "../../include/cvgs_b200.h"
#include "cvgs_device.cuh"

namespace cvgs {

// 6-byte window [p(x1) | p(x1+1)] of one source row, fetched with aligned 32-bit loads.
// Only words that contain at least one needed byte are touched.
struct Window {
    uint32_t lo, hi;  // lo = bytes 0..3, hi = bytes 4..5
};
__device__ __forceinline__ Window load_window(const uint8_t* p, int nbytes) {
    const uintptr_t a = reinterpret_cast<uintptr_t>(p);
    const uint32_t mis = static_cast<uint32_t>(a & 3u);
    const uint32_t* wp = reinterpret_cast<const uint32_t*>(a - mis);
    const uint32_t w0 = __ldg(wp);
    const uint32_t w1 = (mis + nbytes > 4) ? __ldg(wp + 1) : 0u;
    const uint32_t w2 = (mis + nbytes > 8) ? __ldg(wp + 2) : 0u;
    Window w;
    w.lo = __funnelshift_r(w0, w1, mis * 8);
    w.hi = __funnelshift_r(w1, w2, mis * 8);
    return w;
}

// 12-byte window [p(x1) | p(x1+1)] of a row of 16-bit 3-channel pixels (2-byte aligned), same rule.
struct Window12 {
    uint32_t w[3];  // {c0, c1} {c2, c0'} {c1', c2'} as 16-bit halves
};
__device__ __forceinline__ Window12 load_window12(const uint8_t* p, int nbytes) {
    const uintptr_t a = reinterpret_cast<uintptr_t>(p);
    const uint32_t mis = static_cast<uint32_t>(a & 3u);  // 0 or 2
    const uint32_t* wp = reinterpret_cast<const uint32_t*>(a - mis);
    const uint32_t w0 = __ldg(wp);
    const uint32_t w1 = __ldg(wp + 1);  // nbytes >= 6: always needed
    const uint32_t w2 = (mis + nbytes > 8) ? __ldg(wp + 2) : 0u;
    const uint32_t w3 = (mis + nbytes > 12) ? __ldg(wp + 3) : 0u;
    Window12 r;
    r.w[0] = __funnelshift_r(w0, w1, mis * 8);
    r.w[1] = __funnelshift_r(w1, w2, mis * 8);
    r.w[2] = __funnelshift_r(w2, w3, mis * 8);
    return r;
}
// half h (0 low, 1 high) of a word as float: exact for unsigned (mantissa trick) and signed (I2F.S16) samples
template <typename T16>
__device__ __forceinline__ float half_to_f32(uint32_t w, int h) {
    if (sizeof(T16) == 2 && static_cast<T16>(-1) < 0) return static_cast<float>(static_cast<short>(h ? (w >> 16) : (w & 0xffffu)));
    return __fsub_rn(__uint_as_float(__byte_perm(w, 0x4B000000u, h ? 0x7432u : 0x7410u)), 8388608.0f);
}
// CV_16UC3 / CV_16SC3 taps through 32-bit windows: 3-4 loads per row instead of six 16-bit loads.
template <typename T16>
__device__ __forceinline__ void gather_quad_16c3(const PreprocParams& P, const DevCrop& C, int y, int x0, int nvalid,
                                                 float (&v)[4][3]) {
    const AxisTap ty_ = axis_tap(y - C.by1, C.fy);
    const int y2r = min(ty_.i1 + 1, C.h - 1);
    const uint8_t* r0 = C.data + (size_t)ty_.i1 * (size_t)C.pitch;
    const uint8_t* r1 = C.data + (size_t)y2r * (size_t)C.pitch;
#pragma unroll
    for (int p = 0; p < 4; ++p) {
        const int x = x0 + p;
        if (p < nvalid && x >= C.bx1 && x <= C.bx2) {
            const AxisTap tx_ = axis_tap(x - C.bx1, C.fx);
            const bool edge = tx_.i1 + 1 > C.w - 1;  // x2_read == x1
            const int nb = edge ? 6 : 12;
            const Window12 a = load_window12(r0 + 6 * tx_.i1, nb);
            const Window12 b = load_window12(r1 + 6 * tx_.i1, nb);
            const float w00 = __fmul_rn(tx_.w0, ty_.w0), w10 = __fmul_rn(tx_.w1, ty_.w0);
            const float w01 = __fmul_rn(tx_.w0, ty_.w1), w11 = __fmul_rn(tx_.w1, ty_.w1);
            // left taps: halves 0, 1, 2 of the window; right taps: halves 3, 4, 5, or the left pixel again at the edge
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const int hl = c, hr = edge ? c : c + 3;
                const float a0 = half_to_f32<T16>(a.w[hl >> 1], hl & 1), b0 = half_to_f32<T16>(b.w[hl >> 1], hl & 1);
                const float a1 = edge ? a0 : half_to_f32<T16>(a.w[(c + 3) >> 1], (c + 3) & 1);
                const float b1 = edge ? b0 : half_to_f32<T16>(b.w[(c + 3) >> 1], (c + 3) & 1);
                (void)hr;
                v[p][c] = bilerp(a0, a1, b0, b1, w00, w10, w01, w11);
            }
        }
    }
}

template <int NC>
__device__ __forceinline__ void fill_background(const PreprocParams& P, float (&v)[4][NC]) {
#pragma unroll
    for (int p = 0; p < 4; ++p)
#pragma unroll
        for (int c = 0; c < NC; ++c) v[p][c] = P.bg[c];
}

// Resize::exec + Interpolate::exec for output pixels (x0..x0+3, y) of crop C (reference
// resize.cuh:70-82,178-189, interpolation.cuh:57-92).  Values outside the aspect-ratio band are
// the background.  The op chain is NOT applied here.
// 16-bit sources (CV_16UC3 / CV_16SC3): the same arithmetic on ushort3 / short3 taps (the reference promotes them to
// float exactly, core/utils/cuda_vector_utils.h), one 16-bit load per tap channel.
// Also the 4-channel sources (T = unsigned char / unsigned short / short, NC = 4).
template <typename T16, int NC>
__device__ __forceinline__ void gather_quad16(const PreprocParams& P, const DevCrop& C, int y, int x0, int nvalid,
                                              float (&v)[4][NC]) {
    const AxisTap ty_ = axis_tap(y - C.by1, C.fy);
    const int y2r = min(ty_.i1 + 1, C.h - 1);
    const T16* r0 = reinterpret_cast<const T16*>(C.data + (size_t)ty_.i1 * (size_t)C.pitch);
    const T16* r1 = reinterpret_cast<const T16*>(C.data + (size_t)y2r * (size_t)C.pitch);
#pragma unroll
    for (int p = 0; p < 4; ++p) {
        const int x = x0 + p;
        if (p < nvalid && x >= C.bx1 && x <= C.bx2) {
            const AxisTap tx_ = axis_tap(x - C.bx1, C.fx);
            const int x1 = tx_.i1, x2r = min(tx_.i1 + 1, C.w - 1);
            const float w00 = __fmul_rn(tx_.w0, ty_.w0), w10 = __fmul_rn(tx_.w1, ty_.w0);
            const float w01 = __fmul_rn(tx_.w0, ty_.w1), w11 = __fmul_rn(tx_.w1, ty_.w1);
#pragma unroll
            for (int c = 0; c < NC; ++c)
                v[p][c] = bilerp((float)__ldg(r0 + NC * x1 + c), (float)__ldg(r0 + NC * x2r + c), (float)__ldg(r1 + NC * x1 + c),
                                 (float)__ldg(r1 + NC * x2r + c), w00, w10, w01, w11);
        }
    }
}

// One source pixel of an NV12 frame as float RGB: fk::ReadYUV<NV12> + fk::ConvertYUVToRGB<NV12, ., ., false, float3>
// (reference color_conversion.cuh:235-291,296-316).  Rounding sequence of the reference's SASS: per channel
// FMUL(y * m0), FFMA(u, m1, .), FFMA(v, m2, .) -- zero coefficients included -- after y - 16 (bt601 only), u - 128, v - 128.
// The other readers (ReadYUV<NV21 / P010 / P210 / Y210>, :296-345) differ in where the three samples sit; the 10-bit
// formats shift them down by 6, convert in the 10-bit range and multiply the float RGB by 64 afterwards (:226-232,270-291).
__device__ __forceinline__ void nv12_px(const PreprocParams& P, const DevCrop& C, int x, int y, float (&rgb)[3]) {
    float fy, fu, fv;
    const size_t pitch = (size_t)C.pitch;
    if (P.src_type == CVGS_NV12 || P.src_type == CVGS_NV21) {
        const uint8_t* uvp = C.data + pitch * (size_t)C.h + (size_t)(y >> 1) * pitch + 2 * (size_t)(x >> 1);
        fy = (float)__ldg(C.data + (size_t)y * pitch + x);
        const float a = (float)__ldg(uvp), b = (float)__ldg(uvp + 1);
        fu = P.src_type == CVGS_NV12 ? a : b;
        fv = P.src_type == CVGS_NV12 ? b : a;
    } else if (P.src_type == CVGS_Y210) {
        const ushort4 q = __ldg(reinterpret_cast<const ushort4*>(C.data + (size_t)y * pitch) + (x >> 1));
        fy = (float)(((x & 1) ? q.z : q.x) >> 6);
        fu = (float)(q.y >> 6);
        fv = (float)(q.w >> 6);
    } else {  // P010 (4:2:0) / P210 (4:2:2)
        const int cy = P.src_type == CVGS_P010 ? (y >> 1) : y;
        const ushort2 uv = __ldg(reinterpret_cast<const ushort2*>(C.data + pitch * (size_t)C.h + (size_t)cy * pitch) + (x >> 1));
        fy = (float)(__ldg(reinterpret_cast<const unsigned short*>(C.data + (size_t)y * pitch) + x) >> 6);
        fu = (float)(uv.x >> 6);
        fv = (float)(uv.y >> 6);
    }
    const float yy = __fsub_rn(fy, P.yuv[9]);
    const float u = __fsub_rn(fu, P.yuv[10]), v = __fsub_rn(fv, P.yuv[10]);
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        float t = __fmul_rn(yy, P.yuv[3 * r]);
        t = __fmaf_rn(u, P.yuv[3 * r + 1], t);
        rgb[r] = __fmul_rn(__fmaf_rn(v, P.yuv[3 * r + 2], t), P.yuv[11]);
    }
}
__device__ __forceinline__ void gather_quad_nv12(const PreprocParams& P, const DevCrop& C, int y, int x0, int nvalid,
                                                 float (&v)[4][3]) {
    const AxisTap ty_ = axis_tap(y - C.by1, C.fy);
    const int y1 = ty_.i1, y2r = min(ty_.i1 + 1, C.h - 1);
#pragma unroll
    for (int p = 0; p < 4; ++p) {
        const int x = x0 + p;
        if (p < nvalid && x >= C.bx1 && x <= C.bx2) {
            const AxisTap tx_ = axis_tap(x - C.bx1, C.fx);
            const int x1 = tx_.i1, x2r = min(tx_.i1 + 1, C.w - 1);
            const float w00 = __fmul_rn(tx_.w0, ty_.w0), w10 = __fmul_rn(tx_.w1, ty_.w0);
            const float w01 = __fmul_rn(tx_.w0, ty_.w1), w11 = __fmul_rn(tx_.w1, ty_.w1);
            float p00[3], p10[3], p01[3], p11[3];
            nv12_px(P, C, x1, y1, p00);
            nv12_px(P, C, x2r, y1, p10);
            nv12_px(P, C, x1, y2r, p01);
            nv12_px(P, C, x2r, y2r, p11);
#pragma unroll
            for (int c = 0; c < 3; ++c) v[p][c] = bilerp(p00[c], p10[c], p01[c], p11[c], w00, w10, w01, w11);
        }
    }
}

// CV_8UC4 taps as one aligned 32-bit load each (uchar4 rows are 4-byte aligned whenever base and pitch are).
__device__ __forceinline__ void gather_quad_u8c4(const PreprocParams& P, const DevCrop& C, int y, int x0, int nvalid,
                                                 float (&v)[4][4]) {
    const AxisTap ty_ = axis_tap(y - C.by1, C.fy);
    const int y2r = min(ty_.i1 + 1, C.h - 1);
    const uint32_t* r0 = reinterpret_cast<const uint32_t*>(C.data + (size_t)ty_.i1 * (size_t)C.pitch);
    const uint32_t* r1 = reinterpret_cast<const uint32_t*>(C.data + (size_t)y2r * (size_t)C.pitch);
#pragma unroll
    for (int p = 0; p < 4; ++p) {
        const int x = x0 + p;
        if (p < nvalid && x >= C.bx1 && x <= C.bx2) {
            const AxisTap tx_ = axis_tap(x - C.bx1, C.fx);
            const int x1 = tx_.i1, x2r = min(tx_.i1 + 1, C.w - 1);
            const float w00 = __fmul_rn(tx_.w0, ty_.w0), w10 = __fmul_rn(tx_.w1, ty_.w0);
            const float w01 = __fmul_rn(tx_.w0, ty_.w1), w11 = __fmul_rn(tx_.w1, ty_.w1);
            const uint32_t a0 = __ldg(r0 + x1), a1 = __ldg(r0 + x2r), b0 = __ldg(r1 + x1), b1 = __ldg(r1 + x2r);
#pragma unroll
            for (int c = 0; c < 4; ++c)
                v[p][c] = bilerp(u8_to_f32(a0, c), u8_to_f32(a1, c), u8_to_f32(b0, c), u8_to_f32(b1, c), w00, w10, w01, w11);
        }
    }
}

// CV_16UC4 / CV_16SC4 taps as one aligned 64-bit load each (ushort4 / short4 rows are 8-byte aligned whenever base and
// pitch are).
template <typename T16>
__device__ __forceinline__ void gather_quad_16c4(const PreprocParams& P, const DevCrop& C, int y, int x0, int nvalid,
                                                 float (&v)[4][4]) {
    const AxisTap ty_ = axis_tap(y - C.by1, C.fy);
    const int y2r = min(ty_.i1 + 1, C.h - 1);
    const uint2* r0 = reinterpret_cast<const uint2*>(C.data + (size_t)ty_.i1 * (size_t)C.pitch);
    const uint2* r1 = reinterpret_cast<const uint2*>(C.data + (size_t)y2r * (size_t)C.pitch);
#pragma unroll
    for (int p = 0; p < 4; ++p) {
        const int x = x0 + p;
        if (p < nvalid && x >= C.bx1 && x <= C.bx2) {
            const AxisTap tx_ = axis_tap(x - C.bx1, C.fx);
            const int x1 = tx_.i1, x2r = min(tx_.i1 + 1, C.w - 1);
            const float w00 = __fmul_rn(tx_.w0, ty_.w0), w10 = __fmul_rn(tx_.w1, ty_.w0);
            const float w01 = __fmul_rn(tx_.w0, ty_.w1), w11 = __fmul_rn(tx_.w1, ty_.w1);
            const uint2 a0 = __ldg(r0 + x1), a1 = __ldg(r0 + x2r), b0 = __ldg(r1 + x1), b1 = __ldg(r1 + x2r);
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const int h = c & 1;
                v[p][c] = bilerp(half_to_f32<T16>(c < 2 ? a0.x : a0.y, h), half_to_f32<T16>(c < 2 ? a1.x : a1.y, h),
                                 half_to_f32<T16>(c < 2 ? b0.x : b0.y, h), half_to_f32<T16>(c < 2 ? b1.x : b1.y, h), w00, w10, w01, w11);
            }
        }
    }
}

// 4-channel sources (CV_8UC4 / CV_16UC4 / CV_16SC4): same arithmetic on four channels.
__device__ __forceinline__ void gather_quad(const PreprocParams& P, const DevCrop& C, int y, int x0, int nvalid,
                                            float (&v)[4][4]) {
    fill_background<4>(P, v);
    const bool row_in = !P.band_test || (y >= C.by1 && y <= C.by2);
    if (!row_in) return;
    if (P.src_type == CVGS_8UC4) {
        if (((reinterpret_cast<uintptr_t>(C.data) | static_cast<uintptr_t>(C.pitch)) & 3) == 0) gather_quad_u8c4(P, C, y, x0, nvalid, v);
        else gather_quad16<unsigned char, 4>(P, C, y, x0, nvalid, v);
    }
    else {
        const bool aligned8 = ((reinterpret_cast<uintptr_t>(C.data) | static_cast<uintptr_t>(C.pitch)) & 7) == 0;
        if (P.src_type == CVGS_16UC4) {
            if (aligned8) gather_quad_16c4<unsigned short>(P, C, y, x0, nvalid, v);
            else gather_quad16<unsigned short, 4>(P, C, y, x0, nvalid, v);
        } else {
            if (aligned8) gather_quad_16c4<short>(P, C, y, x0, nvalid, v);
            else gather_quad16<short, 4>(P, C, y, x0, nvalid, v);
        }
    }
}

__device__ __forceinline__ void gather_quad(const PreprocParams& P, const DevCrop& C, int y, int x0, int nvalid,
                                            float (&v)[4][3]) {
    fill_background<3>(P, v);
    const bool row_in = !P.band_test || (y >= C.by1 && y <= C.by2);
    if (!row_in) return;
    if (P.src_type != CVGS_8UC3) {
        if (P.src_type == CVGS_16UC3) gather_quad_16c3<unsigned short>(P, C, y, x0, nvalid, v);
        else if (P.src_type == CVGS_16SC3) gather_quad_16c3<short>(P, C, y, x0, nvalid, v);
        else gather_quad_nv12(P, C, y, x0, nvalid, v);
        return;
    }
    const AxisTap ty_ = axis_tap(y - C.by1, C.fy);
    const int y2r = min(ty_.i1 + 1, C.h - 1);
    const uint8_t* r0 = C.data + (size_t)ty_.i1 * (size_t)C.pitch;
    const uint8_t* r1 = C.data + (size_t)y2r * (size_t)C.pitch;
#pragma unroll
    for (int p = 0; p < 4; ++p) {
        const int x = x0 + p;
        if (p < nvalid && x >= C.bx1 && x <= C.bx2) {
            const AxisTap tx_ = axis_tap(x - C.bx1, C.fx);
            const bool edge = tx_.i1 + 1 > C.w - 1;  // x2_read == x1
            const int nb = edge ? 3 : 6;
            const Window a = load_window(r0 + 3 * tx_.i1, nb);
            const Window b = load_window(r1 + 3 * tx_.i1, nb);
            const float w00 = __fmul_rn(tx_.w0, ty_.w0), w10 = __fmul_rn(tx_.w1, ty_.w0);
            const float w01 = __fmul_rn(tx_.w0, ty_.w1), w11 = __fmul_rn(tx_.w1, ty_.w1);
            // right taps: bytes 3,4,5 of the window, or the left pixel again at the edge
            const uint32_t a1 = edge ? a.lo : __funnelshift_r(a.lo, a.hi, 24);
            const uint32_t b1 = edge ? b.lo : __funnelshift_r(b.lo, b.hi, 24);
#pragma unroll
            for (int c = 0; c < 3; ++c)
                v[p][c] = bilerp(u8_to_f32(a.lo, c), u8_to_f32(a1, c), u8_to_f32(b.lo, c), u8_to_f32(b1, c), w00,
                                 w10, w01, w11);
        }
    }
}

}  // namespace cvgs
