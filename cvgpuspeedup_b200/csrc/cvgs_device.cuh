// cvgs_device.cuh -- device-side data structures and arithmetic shared by the sm_100a kernels.
//
// Arithmetic contract (DESIGN.md "Parity"): every float operation is written with an explicit
// rounding intrinsic so that nvcc can neither contract nor reassociate it.  The sequences are
// the ones the reference's fused kernel executes (reference
// fkl/include/fused_kernel/algorithms/image_processing/interpolation.cuh:57-92,
// resize.cuh:178-189, basic_ops/arithmetic.cuh:43-68), pinned from its sm_100a SASS.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace cvgs {

// One source crop plus the resize geometry the host derived for it
// (= fk::RawPtr<_2D,uchar3> + fk::ResizeReadParams, reference resize.cuh:43-58).
struct __align__(16) DevCrop {
    union {
        const uint8_t* data;  // first pixel of the ROI (direct-gather and CircularTensor kernels)
        struct {              // TMA-staged kernel: where the ROI sits inside the tensor map it is staged through
            int32_t xb;       //   byte offset of the ROI's first pixel within a row of the map
            int32_t y0;       //   row of the map that holds the ROI's first row
        } m;
    };
    int32_t w, h;         // source size in pixels
    int32_t pitch;        // bytes between rows
    float fx, fy;         // src_conv_factors
    int32_t bx1, by1;     // band that receives the image (aspect-ratio modes); whole plane otherwise
    int32_t bx2, by2;
    int32_t pad;          // TMA-staged kernel: bits 0..15 staged row bytes, bits 16..31 index of the tensor map
};
static_assert(sizeof(DevCrop) == 48, "DevCrop layout");

// Normalised op chain.  Channel reorders are folded into out_perm, SUB is ADD of the negated
// constant (exact), and under the reference-fused contract MUL followed by ADD/SUB is one FMA.
// Constants are indexed by SOURCE channel ("register space").
// DOP_SET: register c takes a[c] where b[c] != 0 (AddOpaqueAlpha; hoisted to the front of the program by the host).
// DOP_GRAY: register 0 = float(int_rn(0.299 x + 0.587 y + 0.114 z)) with x, y, z = registers (kind >> 8) & 3, (kind >> 12) & 3,
// (kind >> 16) & 3; bit 20: every product and sum rounded on its own (CVGS_FP_SEPARATE) instead of FMUL, FFMA, FFMA;
// bit 21: the stand-alone FMUL is y * 0.587 (else x * 0.299).
// DOP_DIVC: x / d as RN(x * a[c] + RN(x * b[c])) with a + b = 1 / d split by the host (div_const.cpp), used where the host
// has proven the two operations equal to IEEE division for this divisor and this chain (specialize_division).
enum DevOpKind : int32_t { DOP_MUL = 1, DOP_ADD = 2, DOP_DIV = 3, DOP_FMA = 4, DOP_SET = 5, DOP_GRAY = 6, DOP_DIVC = 7 };
struct DevOp {
    int32_t kind;
    float a[4];   // 3-channel launches use [0..2]
    float b[4];
};
struct DevProgram {
    int32_t n_ops;
    int32_t round_u8;     // RoundKind: CVGS_INTERP_ROUND_U8 for the source depth of the launch
    int32_t dst_chan[4];  // register (= source channel) r is written to output channel dst_chan[r]; < 0: not written
    int32_t nc_out;       // channels of the output pixel (differs from the source's after ADD_ALPHA / DROP_ALPHA / GRAY)
    int32_t nregs;        // registers per pixel the chain needs: source channels, or 4 when a 3-channel source gets an alpha
    int32_t special;      // 1: the program has DOP_SET / DOP_GRAY or nc_out != source channels (direct-gather kernel only)
    DevOp ops[8];
};

struct DevPlane {       // one destination image of the per-plane output form (fk::SplitWrite)
    float* data;
    long long pitch;    // floats between rows
};
constexpr long long kNoStore = -0x7fffffffffffffffLL - 1;
struct OutDesc {
    const DevPlane* planes;  // device table [z][source channel], or nullptr for the tensor layouts below
    float* base;
    long long z_stride;  // floats between batch planes
    long long c_stride;  // floats between colour planes
    int32_t px_stride;   // floats between x-adjacent pixels (1 planar, 3 packed)
    int32_t vec4;        // 1: planar rows are 16-byte aligned and W % 4 == 0
    int32_t u8;          // packed 8-bit output after the chain, strides below in bytes: 1 SaturateCast<float, uchar>,
                         // 2 fk::Cast (static_cast: truncation, low byte)
    long long row_pitch; // u8 output: bytes between rows
    long long c_off[4];   // float tensor layouts: register r goes to row + c_off[r] (= dst_chan[r] * c_stride); kNoStore: dropped
    long long row_stride; // float output (tensor layouts): floats between rows (W * px_stride unless the caller set out_row_pitch)
};

struct PreprocParams {
    const DevCrop* crops;  // device table (nullptr when the table rides in the kernel params)
    int32_t n_planes, used;
    int32_t W, H;          // destination size
    int32_t band_test;     // 1 for the aspect-ratio preserving modes
    int32_t src_type;      // CVGS_8UC3 / CVGS_16UC3 / CVGS_16SC3 / CVGS_8UC4 / CVGS_16UC4 / CVGS_16SC4
    int32_t nc;            // channels: 3 or 4
    float yuv[12];         // YUV sources: YCbCr -> RGB matrix (row-major 3x3), [9] luma offset (bt601: 16, 10-bit 64; else 0),
                           // [10] chroma offset (128 / 512), [11] scale back to the stored range (1 / 64)
    float bg[4];           // background / default value (source channel order)
    DevProgram prog;
    OutDesc out;
};

constexpr int kParamCrops = 64;  // batches up to this size carry their descriptors as kernel params
struct ParamCropTable {
    DevCrop c[kParamCrops];
};

// ---------------------------------------------------------------------------------------------
// exact u8 -> f32: byte k of `w` is placed in the mantissa of 2^23, then 2^23 is subtracted.
// One PRMT (alu pipe) + one FADD (fma pipe) instead of a quarter-rate I2F.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float u8_to_f32(uint32_t w, uint32_t k) {
    return __fsub_rn(__uint_as_float(__byte_perm(w, 0x4B000000u, 0x7540u | k)), 8388608.0f);
}

// Bilinear weights of one axis: s = i*f, i1 = floor(s), w1 = s - i1, w0 = (i1+1) - s
// (reference resize.cuh:178-189 + interpolation.cuh:57-70,88-91).
struct AxisTap {
    int32_t i1;
    float w0, w1;
};
__device__ __forceinline__ AxisTap axis_tap(int i, float f) {
    const float s = __fmul_rn((float)i, f);
    AxisTap t;
    t.i1 = __float2int_rd(s);
    const float fl = (float)t.i1;
    t.w1 = __fsub_rn(s, fl);
    t.w0 = __fsub_rn(__fadd_rn(fl, 1.0f), s);  // (float)(i1+1) == fl + 1 exactly for |i1| < 2^24
    return t;
}

// One channel of Interpolate<INTER_LINEAR>::exec in the order nvcc emits for the reference:
// FMUL(p10*w10), FFMA(p00,w00), FFMA(p01,w01), FFMA(p11,w11).
__device__ __forceinline__ float bilerp(float p00, float p10, float p01, float p11, float w00, float w10,
                                        float w01, float w11) {
    float t = __fmul_rn(p10, w10);
    t = __fmaf_rn(p00, w00, t);
    t = __fmaf_rn(p01, w01, t);
    t = __fmaf_rn(p11, w11, t);
    return t;
}

// SaturateCast<float, uchar> then back to float (reference saturate.cuh:127-147).
__device__ __forceinline__ float round_sat_u8(float v) {
    const unsigned u = __float2uint_rn(v);
    return (float)(u > 255u ? 255u : u);
}
// DevProgram::round_u8 codes: the interpolated value is rounded and saturated to the range of the source depth
enum RoundKind : int32_t { ROUND_NONE = 0, ROUND_U8 = 1, ROUND_U16 = 2, ROUND_S16 = 3 };
// SaturateCast<float, ushort> (saturate.cuh:267-298) / <float, short> (:358-378) then back to float.
__device__ __forceinline__ float round_sat_kind(float v, int kind) {
    if (kind == ROUND_U16) {
        const unsigned u = __float2uint_rn(v);
        return (float)(u > 65535u ? 65535u : u);
    }
    if (kind == ROUND_S16) {
        const int i = __float2int_rn(v);
        return (float)(i > 32767 ? 32767 : (i < -32768 ? -32768 : i));
    }
    return round_sat_u8(v);
}

// Apply the normalised chain to N values laid out as v[pixel][channel].
template <int NPIX, int NC = 3>
__device__ __forceinline__ void apply_program(const DevProgram& prog, float (&v)[NPIX][NC]) {
    if (prog.round_u8) {
#pragma unroll
        for (int p = 0; p < NPIX; ++p)
#pragma unroll
            for (int c = 0; c < NC; ++c) v[p][c] = round_sat_kind(v[p][c], prog.round_u8);
    }
    for (int i = 0; i < prog.n_ops; ++i) {
        const DevOp& op = prog.ops[i];
        switch (op.kind) {
            case DOP_FMA:
#pragma unroll
                for (int p = 0; p < NPIX; ++p)
#pragma unroll
                    for (int c = 0; c < NC; ++c) v[p][c] = __fmaf_rn(v[p][c], op.a[c], op.b[c]);
                break;
            case DOP_MUL:
#pragma unroll
                for (int p = 0; p < NPIX; ++p)
#pragma unroll
                    for (int c = 0; c < NC; ++c) v[p][c] = __fmul_rn(v[p][c], op.a[c]);
                break;
            case DOP_ADD:
#pragma unroll
                for (int p = 0; p < NPIX; ++p)
#pragma unroll
                    for (int c = 0; c < NC; ++c) v[p][c] = __fadd_rn(v[p][c], op.a[c]);
                break;
            case DOP_DIV:
#pragma unroll
                for (int p = 0; p < NPIX; ++p)
#pragma unroll
                    for (int c = 0; c < NC; ++c) v[p][c] = __fdiv_rn(v[p][c], op.a[c]);
                break;
            case DOP_DIVC:
#pragma unroll
                for (int p = 0; p < NPIX; ++p)
#pragma unroll
                    for (int c = 0; c < NC; ++c) v[p][c] = __fmaf_rn(v[p][c], op.a[c], __fmul_rn(v[p][c], op.b[c]));
                break;
            case DOP_SET:
#pragma unroll
                for (int p = 0; p < NPIX; ++p)
#pragma unroll
                    for (int c = 0; c < NC; ++c) v[p][c] = op.b[c] != 0.f ? op.a[c] : v[p][c];
                break;
            default:
                if ((op.kind & 0xff) == DOP_GRAY) {
                    const int rx = (op.kind >> 8) & 3, ry = (op.kind >> 12) & 3, rz = (op.kind >> 16) & 3;
                    const bool separate = (op.kind >> 20) & 1, y_first = (op.kind >> 21) & 1;
#pragma unroll
                    for (int p = 0; p < NPIX; ++p) {
                        // selects instead of dynamic indexing: v stays in registers
                        auto pick = [&](int r) {
                            float t = v[p][0];
#pragma unroll
                            for (int c = 1; c < NC; ++c) t = r == c ? v[p][c] : t;
                            return t;
                        };
                        const float x = pick(rx), y = pick(ry), z = pick(rz);
                        float t;
                        if (separate) {
                            t = __fadd_rn(__fadd_rn(__fmul_rn(x, 0.299f), __fmul_rn(y, 0.587f)), __fmul_rn(z, 0.114f));
                        } else if (y_first) {
                            t = __fmaf_rn(z, 0.114f, __fmaf_rn(x, 0.299f, __fmul_rn(y, 0.587f)));
                        } else {
                            t = __fmaf_rn(z, 0.114f, __fmaf_rn(y, 0.587f, __fmul_rn(x, 0.299f)));
                        }
                        // the reference's RGB2Gray<I, float> rounds the luminance to an integer (its is_signed branch)
                        v[p][0] = static_cast<float>(__float2int_rn(t));
                    }
                }
                break;
        }
    }
}

__device__ __forceinline__ void st_cs_f32x4(float* p, float a, float b, float c, float d) {
    asm volatile("st.global.cs.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d)
                 : "memory");
}
// predicated store (no branch: keeps the surrounding code in one basic block)
#ifndef CVGS_ST_F32  // diagnostic builds try other cache policies (scripts/diag_build.sh)
#define CVGS_ST_F32 "st.global.cs.f32"
#endif
__device__ __forceinline__ void st_cs_f32_if(bool pred, float* p, float a) {
    asm volatile("{\n.reg .pred q;\nsetp.ne.u32 q, %0, 0;\n@q " CVGS_ST_F32 " [%1], %2;\n}" ::"r"((unsigned)pred), "l"(p), "f"(a) : "memory");
}
__device__ __forceinline__ void st_cs_f32(float* p, float a) {
    asm volatile(CVGS_ST_F32 " [%0], %1;" ::"l"(p), "f"(a) : "memory");
}

// Store NPIX x-adjacent pixels (first one at column x) of row y, plane z.
template <int NPIX, int NC = 3>
__device__ __forceinline__ void store_pixels(const PreprocParams& P, int z, int y, int x, int nvalid,
                                             const float (&v)[NPIX][NC]) {
    const OutDesc& o = P.out;
    if (o.u8) {  // convertTo<CV_32FCn, CV_8UCn> + packed write: byte dst_chan[r] of pixel p is register r
        const int nco = P.prog.nc_out;  // channels of the written pixel (the chain may have added / dropped the alpha)
        uint8_t* px = reinterpret_cast<uint8_t*>(o.base) + (long long)z * o.z_stride + (long long)y * o.row_pitch + (long long)nco * x;
        if (NPIX == 4 && nvalid == 4 && nco == NC && o.vec4) {
            // the four pixels of the thread are NC * 4 contiguous bytes starting at a multiple of 4 (x % 4 == 0, rows and
            // base 4-byte aligned: vec4): assemble them in registers and store NC words instead of 4 * NC bytes
            uint32_t b[4][NC];  // [pixel][output channel]
#pragma unroll
            for (int p = 0; p < 4; ++p)
#pragma unroll
                for (int c = 0; c < NC; ++c) {
                    float t = v[p][0];
#pragma unroll
                    for (int r = 1; r < NC; ++r) t = P.prog.dst_chan[r] == c ? v[p][r] : t;  // uniform selects
                    b[p][c] = o.u8 == 2 ? (__float2uint_rz(t) & 0xffu) : (uint32_t)round_sat_u8(t);
                }
            uint32_t* wp = reinterpret_cast<uint32_t*>(px);
#pragma unroll
            for (int k = 0; k < NC; ++k) {  // word k holds bytes 4k .. 4k+3 of the NC * 4
                uint32_t w = 0;
#pragma unroll
                for (int j = 0; j < 4; ++j) w |= b[(4 * k + j) / NC][(4 * k + j) % NC] << (8 * j);
                asm volatile("st.global.cs.u32 [%0], %1;" ::"l"(wp + k), "r"(w) : "memory");
            }
            return;
        }
#pragma unroll
        for (int p = 0; p < NPIX; ++p)
            if (p < nvalid) {
#pragma unroll
                for (int r = 0; r < NC; ++r)
                    if (P.prog.dst_chan[r] >= 0)
                        px[nco * p + P.prog.dst_chan[r]] =
                            o.u8 == 2 ? (uint8_t)__float2uint_rz(v[p][r]) : (uint8_t)round_sat_u8(v[p][r]);
            }
        return;
    }
    if (o.planes) {  // fk::SplitWrite: a table of destination images indexed by SOURCE channel (the host applied dst_chan)
#pragma unroll
        for (int r = 0; r < NC; ++r) {
            const DevPlane pl = o.planes[z * NC + r];
            float* dst = pl.data + (long long)y * pl.pitch + x;
#pragma unroll
            for (int p = 0; p < NPIX; ++p)
                if (p < nvalid) st_cs_f32(dst + p, v[p][r]);
        }
        return;
    }
    // the channel reorder costs nothing: it only changes which plane register r goes to (c_off, computed on the host)
    float* row = o.base + (long long)z * o.z_stride + (long long)y * o.row_stride + (long long)x * o.px_stride;
    if (NPIX == 4 && o.vec4 && nvalid == 4) {
#pragma unroll
        for (int r = 0; r < NC; ++r)
            if (o.c_off[r] != kNoStore) st_cs_f32x4(row + o.c_off[r], v[0][r], v[1][r], v[2][r], v[3][r]);
        return;
    }
#pragma unroll
    for (int r = 0; r < NC; ++r) {
        if (o.c_off[r] == kNoStore) continue;  // register dropped by a channel-count changing conversion
        float* dst = row + o.c_off[r];
#pragma unroll
        for (int p = 0; p < NPIX; ++p)
            if (p < nvalid) st_cs_f32(dst + (long long)p * o.px_stride, v[p][r]);
    }
}

}  // namespace cvgs
