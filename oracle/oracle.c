/*
 * oracle/oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement of the reference's algorithm for the hot path.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may load this library, and only as the checker / CPU baseline -- never as a fallback of
 * the product path (cvgpuspeedup_b200 fails loudly when its CUDA library is missing).
 *
 * Pinning: the arithmetic below was derived from the reference sources cited on each
 * function AND from the SASS nvcc 12.9 emits for the reference's own fused kernel
 * (oracle/ref_harness.cu compiled for sm_100a): interpolation = FMUL(p10*w10) then
 * FFMA(p00,w00), FFMA(p01,w01), FFMA(p11,w11); Mul->Sub contracted to one FFMA; IEEE division.
 * tests/test_ref_parity.py compares this file bit-for-bit with the reference kernel itself
 * (oracle/_ref/libfkref_*.so) on a B200, and tests/test_oracle_golden.py checks it against the
 * known answers the reference's tests hold for this path.
 *
 * Build: gcc -O2 -ffp-contract=off -mfma -fopenmp -shared -fPIC (see oracle/Makefile).
 * -ffp-contract=off + explicit fmaf(): every rounding below is written out.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../include/cvgs_b200.h"

#ifdef _OPENMP
#include <omp.h>
#endif

/* Geometry of one crop's resize, as stored in fk::ResizeReadParams
 * (reference fkl/.../image_processing/resize.cuh:43-58). */
typedef struct oracle_geom {
    float fx, fy;          /* src_conv_factors                                  */
    int x1, y1, x2, y2;    /* band that receives the resized image (AR modes)   */
} oracle_geom_t;

/* cxp::round, reference fkl/.../constexpr_libs/constexpr_cmath.cuh:37-48 (float instance). */
static float cxp_roundf(float x) {
    if (x != x || (x == x && x != 0.0f && x + x == x)) return x;
    return (x > 0.0f) ? (float)(int)(x + 0.5f) : (float)(int)(x - 0.5f);
}

/* Resize::build (IGNORE_AR) resize.cuh:100-114; AR variants :116-161; compute_target_size :191-216. */
void oracle_resize_geometry(int src_w, int src_h, int dst_w, int dst_h, int aspect_mode,
                            oracle_geom_t* g) {
    if (aspect_mode == CVGS_IGNORE_AR) {
        const double cfx = (double)dst_w / (double)src_w;
        const double cfy = (double)dst_h / (double)src_h;
        g->fx = (float)(1.0 / cfx);
        g->fy = (float)(1.0 / cfy);
        g->x1 = 0; g->y1 = 0; g->x2 = dst_w - 1; g->y2 = dst_h - 1;
        return;
    }
    /* compute_target_size */
    int tw, th;
    {
        const float scaleFactor = dst_h / (float)src_h;
        const int targetHeight = dst_h;
        const int targetWidth = (int)cxp_roundf(scaleFactor * src_w);
        if (aspect_mode == CVGS_PRESERVE_AR_RN_EVEN) {
            const int targetWidthTemp = targetWidth - (targetWidth % 2);
            if (targetWidthTemp > dst_w) {
                const float scaleFactorTemp = dst_w / (float)src_w;
                const int targetHeightTemp = (int)cxp_roundf(scaleFactorTemp * src_h);
                tw = dst_w; th = targetHeightTemp - (targetHeightTemp % 2);
            } else {
                tw = targetWidthTemp; th = targetHeight;
            }
        } else {
            if (targetWidth > dst_w) {
                const float scaleFactorTemp = dst_w / (float)src_w;
                tw = dst_w; th = (int)cxp_roundf(scaleFactorTemp * src_h);
            } else {
                tw = targetWidth; th = targetHeight;
            }
        }
    }
    const double cfx = (double)tw / src_w;
    const double cfy = (double)th / src_h;
    g->fx = (float)(1.0 / cfx);
    g->fy = (float)(1.0 / cfy);
    g->x1 = (aspect_mode == CVGS_PRESERVE_AR_LEFT) ? 0 : (int)((dst_w - tw) / 2);
    g->y1 = (int)((dst_h - th) / 2);
    g->x2 = g->x1 + tw - 1;
    g->y2 = g->y1 + th - 1;
}

/* SaturateCast<float, uchar> device path: __float2uint_rn then clamp, saturate.cuh:127-147. */
static float round_sat_u8(float v) {
    if (!(v > 0.0f)) return 0.0f;          /* negative and NaN -> 0, as cvt.rni.u32.f32 */
    const float r = nearbyintf(v);         /* default rounding mode = RN-even           */
    return r > 255.0f ? 255.0f : r;
}
/* SaturateCast<float, ushort> (saturate.cuh:267-298) and <float, short> (:358-378): round to nearest even, clamp. */
static float round_sat_u16(float v) {
    if (!(v > 0.0f)) return 0.0f;
    const float r = nearbyintf(v);
    return r > 65535.0f ? 65535.0f : r;
}
static float round_sat_s16(float v) {
    if (v != v) return 0.0f;               /* cvt.rni.s32.f32 of NaN is 0 */
    const float r = nearbyintf(v);
    /* the value passes through a short: (-0.5, -0] comes back as +0, not as the -0 nearbyintf returns */
    return (r > 32767.0f ? 32767.0f : (r < -32768.0f ? -32768.0f : r)) + 0.0f;
}
static float round_sat_src(float v, int src_type) {
    return (src_type == CVGS_16UC3 || src_type == CVGS_16UC4)   ? round_sat_u16(v)
           : (src_type == CVGS_16SC3 || src_type == CVGS_16SC4) ? round_sat_s16(v)
                                                                : round_sat_u8(v);
}
/* channel ch of pixel x of a source row, as the float the reference's uchar3/ushort3/short3 * float promotes it to */
/* fk::ReadYUV<NV12> + fk::ConvertYUVToRGB<NV12, range, primaries, false, float3> for one source pixel
 * (color_conversion.cuh:235-291, matrices :171-214).  Rounding sequence as nvcc compiles MxVFloat3 + VectorReduce for
 * the reference (SASS): per output channel FMUL(y * m0), FFMA(u, m1, .), FFMA(v, m2, .), zero coefficients included;
 * bt601 subtracts 16 from the luma first, every standard subtracts 128 from the chroma. */
static const float kYuvMatrix[4][9] = {
    {1.164383562f, 0.f, 1.596026786f, 1.164383562f, -0.39176229f, -0.812967647f, 1.164383562f, 2.017232143f, 0.f},
    {1.f, 0.f, 1.5748f, 1.f, -0.1873f, -0.4681f, 1.f, 1.8556f, 0.f},
    {1.f, 0.f, 1.402f, 1.f, -0.34414f, -0.71414f, 1.f, 1.772f, 0.f},
    {1.f, 0.f, 1.4746f, 1.f, -0.16455312684366f, -0.57135312684366f, 1.f, 1.8814f, 0.f}};
/* The other readers (ReadYUV<NV21 / P010 / P210 / Y210>, :296-345): NV21 swaps the chroma bytes; the 10-bit formats keep
 * the sample in the high bits of 16-bit words (ShiftRight by shiftFactor<p10bit> = 6, :183-186,278), are converted in the
 * 10-bit range (subCoefficients<p10bit> {64, 512}, :106-111) and the float RGB is scaled back by floatShiftFactor = 64
 * (NormalizeColorRangeDepth, :226-232). */
static void nv12_px(const cvgs_crop_t* c, int src_type, int x, int y, int standard, float rgb[3]) {
    const uint8_t* base = (const uint8_t*)c->data;
    const size_t pitch = (size_t)c->pitch;
    float yy, u, v;
    const int ten = src_type == CVGS_P010 || src_type == CVGS_P210 || src_type == CVGS_Y210;
    if (src_type == CVGS_NV12 || src_type == CVGS_NV21) {
        const uint8_t* uv = base + pitch * (size_t)c->height + (size_t)(y >> 1) * pitch + 2 * (size_t)(x >> 1);
        yy = (float)base[(size_t)y * pitch + x];
        u = (float)uv[src_type == CVGS_NV12 ? 0 : 1];
        v = (float)uv[src_type == CVGS_NV12 ? 1 : 0];
    } else if (src_type == CVGS_Y210) {
        const uint16_t* q = (const uint16_t*)(base + (size_t)y * pitch) + 4 * (size_t)(x >> 1);
        yy = (float)(q[(x & 1) ? 2 : 0] >> 6);
        u = (float)(q[1] >> 6);
        v = (float)(q[3] >> 6);
    } else {
        const int cy = src_type == CVGS_P010 ? (y >> 1) : y;
        const uint16_t* uv = (const uint16_t*)(base + pitch * (size_t)c->height + (size_t)cy * pitch) + 2 * (size_t)(x >> 1);
        yy = (float)(((const uint16_t*)(base + (size_t)y * pitch))[x] >> 6);
        u = (float)(uv[0] >> 6);
        v = (float)(uv[1] >> 6);
    }
    if (standard == CVGS_YUV_BT601_FULL) yy = yy - (ten ? 64.0f : 16.0f);
    u = u - (ten ? 512.0f : 128.0f);
    v = v - (ten ? 512.0f : 128.0f);
    const float* m = kYuvMatrix[standard];
    for (int r = 0; r < 3; ++r) {
        float t = yy * m[3 * r];
        t = fmaf(u, m[3 * r + 1], t);
        t = fmaf(v, m[3 * r + 2], t);
        rgb[r] = ten ? t * 64.0f : t;
    }
}

static int n_channels(int src_type) {
    return (src_type == CVGS_8UC4 || src_type == CVGS_16UC4 || src_type == CVGS_16SC4) ? 4 : 3;
}
static float src_px(const uint8_t* row, int x, int ch, int src_type) {
    const int nc = n_channels(src_type);
    if (src_type == CVGS_16UC3 || src_type == CVGS_16UC4) return (float)((const uint16_t*)row)[nc * x + ch];
    if (src_type == CVGS_16SC3 || src_type == CVGS_16SC4) return (float)((const int16_t*)row)[nc * x + ch];
    return (float)row[nc * x + ch];
}

/* One output pixel of Resize::exec + Interpolate<INTER_LINEAR>::exec
 * (resize.cuh:70-82,178-189; interpolation.cuh:57-92; PerThreadRead ptr_nd.cuh:41-45).
 * Rounding sequence = what nvcc emits for the reference kernel (see file header). */
static void resize_pixel(const cvgs_crop_t* c, const oracle_geom_t* g, int aspect_mode, int src_type, int yuv_standard,
                         int x, int y, const float* bg, float out[4]) {
    const int nc = n_channels(src_type);
    if (aspect_mode != CVGS_IGNORE_AR) {
        if (!(x >= g->x1 && x <= g->x2 && y >= g->y1 && y <= g->y2)) {
            for (int ch = 0; ch < nc; ++ch) out[ch] = bg[ch];
            return;
        }
        x -= g->x1; y -= g->y1;
    }
    const float src_x = (float)x * g->fx;
    const float src_y = (float)y * g->fy;
    const int x1 = (int)floorf(src_x);
    const int y1 = (int)floorf(src_y);
    const int x2 = x1 + 1, y2 = y1 + 1;
    const int x2r = x2 < c->width - 1 ? x2 : c->width - 1;
    const int y2r = y2 < c->height - 1 ? y2 : c->height - 1;
    const float wx1 = src_x - (float)x1, wx0 = (float)x2 - src_x;
    const float wy1 = src_y - (float)y1, wy0 = (float)y2 - src_y;
    const float w00 = wx0 * wy0, w10 = wx1 * wy0, w01 = wx0 * wy1, w11 = wx1 * wy1;
    const uint8_t* base = (const uint8_t*)c->data;
    const uint8_t* r0 = base + (size_t)y1 * (size_t)c->pitch;
    const uint8_t* r1 = base + (size_t)y2r * (size_t)c->pitch;
    if (CVGS_IS_YUV(src_type)) { /* the back-function of the resize is read + colour conversion: taps are float RGB */
        float p00[3], p10[3], p01[3], p11[3];
        nv12_px(c, src_type, x1, y1, yuv_standard, p00);
        nv12_px(c, src_type, x2r, y1, yuv_standard, p10);
        nv12_px(c, src_type, x1, y2r, yuv_standard, p01);
        nv12_px(c, src_type, x2r, y2r, yuv_standard, p11);
        for (int ch = 0; ch < 3; ++ch) {
            float t = p10[ch] * w10;
            t = fmaf(p00[ch], w00, t);
            t = fmaf(p01[ch], w01, t);
            t = fmaf(p11[ch], w11, t);
            out[ch] = t;
        }
        return;
    }
    for (int ch = 0; ch < nc; ++ch) {
        const float p00 = src_px(r0, x1, ch, src_type), p10 = src_px(r0, x2r, ch, src_type);
        const float p01 = src_px(r1, x1, ch, src_type), p11 = src_px(r1, x2r, ch, src_type);
        float t = p10 * w10;
        t = fmaf(p00, w00, t);
        t = fmaf(p01, w01, t);
        t = fmaf(p11, w11, t);
        out[ch] = t;
    }
}

/* One output pixel of fk::Warping<WT, PerThreadRead<_2D, uchar3>> (warping.cuh:43-91): WarpingCoords (:43-63), the
 * bounds test (:78-82) and Interpolate<INTER_LINEAR> at the warped coordinate (interpolation.cuh:57-92).  Rounding
 * sequence of the reference's SASS: FMUL(m01*y), FFMA(m00, x, .), FADD(m02) per row; perspective multiplies by the
 * correctly rounded reciprocal of the third row. */
static void warp_pixel(const cvgs_crop_t* c, const cvgs_warp_t* wp, int src_type, int x, int y, float out[4]) {
    const int nc = n_channels(src_type);
    const float fx = (float)x, fy = (float)y;
    const float* m = wp->m;
    float sx = fmaf(m[0], fx, m[1] * fy) + m[2];
    float sy = fmaf(m[3], fx, m[4] * fy) + m[5];
    if (wp->type == CVGS_WARP_PERSPECTIVE) {
        const float coeff = 1.0f / (fmaf(m[6], fx, m[7] * fy) + m[8]);
        sx = coeff * sx;
        sy = coeff * sy;
    }
    if (!(sx >= 0.f && sx < (float)c->width && sy >= 0.f && sy < (float)c->height)) {
        out[0] = out[1] = out[2] = out[3] = 0.f;
        return;
    }
    const int x1 = (int)floorf(sx), y1 = (int)floorf(sy);
    const int x2 = x1 + 1, y2 = y1 + 1;
    const int x2r = x2 < c->width - 1 ? x2 : c->width - 1;
    const int y2r = y2 < c->height - 1 ? y2 : c->height - 1;
    const float wx1 = sx - (float)x1, wx0 = (float)x2 - sx;
    const float wy1 = sy - (float)y1, wy0 = (float)y2 - sy;
    const float w00 = wx0 * wy0, w10 = wx1 * wy0, w01 = wx0 * wy1, w11 = wx1 * wy1;
    const uint8_t* r0 = (const uint8_t*)c->data + (size_t)y1 * (size_t)c->pitch;
    const uint8_t* r1 = (const uint8_t*)c->data + (size_t)y2r * (size_t)c->pitch;
    for (int ch = 0; ch < nc; ++ch) {
        float t = src_px(r0, x2r, ch, src_type) * w10;
        t = fmaf(src_px(r0, x1, ch, src_type), w00, t);
        t = fmaf(src_px(r1, x1, ch, src_type), w01, t);
        t = fmaf(src_px(r1, x2r, ch, src_type), w11, t);
        out[ch] = t;
    }
}

/* Unary/Binary op chain, TransformDPP::operate (data_parallel_patterns.cuh:66-79);
 * Mul/Sub/Div/Add arithmetic.cuh:43-68; VectorReorder cuda_vector.cuh:45-54. */
/* Channels of the pixel the chain ends with (ADD_ALPHA 3 -> 4, DROP_ALPHA 4 -> 3, GRAY -> 1:
 * color_conversion.cuh:364-461). */
static int chain_out_channels(const cvgs_pipeline_t* p) {
    int nc = n_channels(p->src_type);
    for (int i = 0; i < p->n_ops; ++i) {
        if (p->ops[i].kind == CVGS_OP_ADD_ALPHA) nc = 4;
        else if (p->ops[i].kind == CVGS_OP_DROP_ALPHA) nc = 3;
        else if (p->ops[i].kind == CVGS_OP_GRAY) nc = 1;
    }
    return nc;
}

static void apply_chain(const cvgs_pipeline_t* p, float v[4]) {
    int nc = n_channels(p->src_type);
    if (p->interp_mode == CVGS_INTERP_ROUND_U8)
        for (int c = 0; c < nc; ++c) v[c] = round_sat_src(v[c], p->src_type);
    /* nvcc contracts (x*m) -/+ s of the inlined chain into one FMA; channel reorders and the alpha conversions in
     * between are only register renaming and do not prevent it.  mul_x / mul_m remember, per channel, the factors of
     * a product that nothing has consumed yet. */
    const int fused = p->fp_contract == CVGS_FP_REFERENCE_FUSED;
    float mul_x[4] = {0, 0, 0, 0}, mul_m[4] = {0, 0, 0, 0};
    int pending[4] = {0, 0, 0, 0};
    for (int i = 0; i < p->n_ops; ++i) {
        const cvgs_op_t* op = &p->ops[i];
        switch (op->kind) {
            case CVGS_OP_MUL:
                for (int c = 0; c < nc; ++c) {
                    mul_x[c] = v[c];
                    mul_m[c] = op->v[c];
                    pending[c] = fused;
                    v[c] = v[c] * op->v[c];
                }
                break;
            case CVGS_OP_SUB:
            case CVGS_OP_ADD:
                for (int c = 0; c < nc; ++c) {
                    const float a = op->kind == CVGS_OP_SUB ? -op->v[c] : op->v[c];
                    v[c] = pending[c] ? fmaf(mul_x[c], mul_m[c], a) : v[c] + a;
                    pending[c] = 0;
                }
                break;
            case CVGS_OP_DIV:
                for (int c = 0; c < nc; ++c) { v[c] = v[c] / op->v[c]; pending[c] = 0; }
                break;
            case CVGS_OP_REORDER: {
                float t[4], tx[4], tm[4];
                int tp[4];
                for (int c = 0; c < nc; ++c) {
                    const int s = op->perm[c];
                    t[c] = v[s]; tx[c] = mul_x[s]; tm[c] = mul_m[s]; tp[c] = pending[s];
                }
                for (int c = 0; c < nc; ++c) { v[c] = t[c]; mul_x[c] = tx[c]; mul_m[c] = tm[c]; pending[c] = tp[c]; }
                break;
            }
            case CVGS_OP_ADD_ALPHA: /* AddOpaqueAlpha: AddLast(input, alpha), color_conversion.cuh:122-130 */
                v[3] = op->v[0];
                pending[3] = 0;
                nc = 4;
                break;
            case CVGS_OP_DROP_ALPHA: /* Discard<I, VectorType_t<VBase<I>, 3>> */
                nc = 3;
                break;
            case CVGS_OP_GRAY: { /* RGB2Gray<I, float>::compute_luminance, color_conversion.cuh:64-66; as compiled by
                                    nvcc: FMUL(x, 0.299), FFMA(y, 0.587, .), FFMA(z, 0.114, .) */
                /* Which of the first two products stays a stand-alone FMUL is nvcc's choice per instantiation:
                 * RGB2GRAY / RGBA2GRAY multiply y first (op->perm[0] == 1), the codes with the fused reorder in front
                 * (BGR2GRAY / BGRA2GRAY) multiply x first (perm[0] == 0) -- read off the reference's SASS. */
                float t;
                if (!fused) {
                    const float a = v[0] * 0.299f, u = v[1] * 0.587f, w = v[2] * 0.114f;
                    t = (a + u) + w;
                } else if (op->perm[0] == 1) {
                    t = v[1] * 0.587f;
                    t = fmaf(v[0], 0.299f, t);
                    t = fmaf(v[2], 0.114f, t);
                } else {
                    t = v[0] * 0.299f;
                    t = fmaf(v[1], 0.587f, t);
                    t = fmaf(v[2], 0.114f, t);
                }
                /* RGB2Gray<I, float> takes the `std::is_signed_v<OutputType>` branch (true for float,
                 * color_conversion.cuh:55-60): the luminance goes through __float2int_rn -- rounded to the nearest even
                 * integer, saturated to the int range, NaN -> 0 -- and back to float. */
                if (t != t) t = 0.f;
                else {
                    t = nearbyintf(t);
                    t = t >= 2147483648.f ? 2147483648.f : (t < -2147483648.f ? -2147483648.f : (float)(int)t);
                }
                v[0] = t;
                pending[0] = 0;
                nc = 1;
                break;
            }
            default: break;
        }
    }
}

/* Output addressing: TensorSplit memory_operations.cuh:168-188 + PtrAccessor<_3D> ptr_nd.cuh:53-63;
 * TensorTSplit :197-220 + PtrAccessor<T3D> ptr_nd.cuh:65-77; PerThreadWrite<_3D>. */
static void store_pixel(const cvgs_pipeline_t* p, int n_planes, int z, int y, int x, const float v[4]) {
    float* out = (float*)p->out;
    const int nc = chain_out_channels(p);
    const int64_t W = p->dst_width, H = p->dst_height;
    if (p->dst_type == CVGS_8UC3 || p->dst_type == CVGS_8UC4) { /* convertTo<CV_32FCn, CV_8UCn> + PerThreadWrite: SaturateCast saturate.cuh:127-147 */
        const int64_t rp = p->out_row_pitch ? p->out_row_pitch : nc * W;
        const int64_t ps = p->out_plane_stride ? p->out_plane_stride : rp * H;
        uint8_t* b = (uint8_t*)p->out + z * ps + y * rp + nc * x;
        for (int c = 0; c < nc; ++c)  /* u8_cast: fk::Cast = static_cast (truncation, cast.cuh:22-29) */
            b[c] = p->u8_cast ? (uint8_t)(unsigned)v[c] : (uint8_t)round_sat_u8(v[c]);
        return;
    }
    switch (p->out_layout) {
        case CVGS_OUT_NCHW: {
            const int64_t ps = p->out_plane_stride ? p->out_plane_stride : nc * W * H;
            float* b = out + z * ps + y * W + x;
            for (int c = 0; c < nc; ++c) b[c * W * H] = v[c];
            break;
        }
        case CVGS_OUT_CNHW: {
            const int64_t ps = p->out_plane_stride ? p->out_plane_stride : W * H;
            float* b = out + z * ps + y * W + x;
            for (int c = 0; c < nc; ++c) b[c * ps * n_planes] = v[c];
            break;
        }
        case CVGS_OUT_PLANES: { /* SplitWrite memory_operations.cuh:331-360: one RawPtr<_2D,float> per channel */
            const cvgs_plane_t* pl = (const cvgs_plane_t*)p->out + (int64_t)z * nc;
            for (int c = 0; c < nc; ++c)
                *(float*)((char*)pl[c].data + (int64_t)y * pl[c].pitch_bytes + (int64_t)x * 4) = v[c];
            break;
        }
        default: { /* packed pixels; rows may be padded (PerThreadWrite<_2D> into a pitched GpuMat) */
            const int64_t rs = p->out_row_pitch ? p->out_row_pitch / 4 : nc * W;
            const int64_t ps = p->out_plane_stride ? p->out_plane_stride : rs * H;
            float* b = out + z * ps + y * rs + x * nc;
            for (int c = 0; c < nc; ++c) b[c] = v[c];
        }
    }
}

/* The whole fused call: BatchRead<N, CONDITIONAL_WITH_DEFAULT>::exec (batch_operations.cuh:222-229)
 * -> chain -> write, for every (x, y, z) of ActiveThreads{dst_w, dst_h, N}.
 * All pointers are HOST pointers here.  nthreads <= 0 -> all cores. Returns 0 / 1 (bad args). */
int oracle_preproc(const cvgs_crop_t* crops, int n_planes, int used, const cvgs_pipeline_t* p,
                   int nthreads) {
    if (!crops || !p || !p->out || n_planes <= 0 || used < 0 ||
        (p->src_type != CVGS_8UC3 && p->src_type != CVGS_16UC3 && p->src_type != CVGS_16SC3 &&
         p->src_type != CVGS_8UC4 && p->src_type != CVGS_16UC4 && p->src_type != CVGS_16SC4 && !CVGS_IS_YUV(p->src_type)) ||
        (CVGS_IS_YUV(p->src_type) && (p->yuv_standard < 0 || p->yuv_standard > 3)) ||
        p->dst_width <= 0 || p->dst_height <= 0 || p->n_ops < 0 || p->n_ops > CVGS_MAX_OPS)
        return 1;
    if (used > n_planes) used = n_planes;
    oracle_geom_t* geoms = (oracle_geom_t*)calloc((size_t)n_planes, sizeof(oracle_geom_t));
    for (int z = 0; z < used; ++z)
        oracle_resize_geometry(crops[z].width, crops[z].height, p->dst_width, p->dst_height,
                               p->aspect_mode, &geoms[z]);
    const int H = p->dst_height, W = p->dst_width;
    const long total_rows = (long)n_planes * H;
#ifdef _OPENMP
    if (nthreads <= 0) nthreads = omp_get_max_threads();
#pragma omp parallel for schedule(dynamic, 8) num_threads(nthreads)
#endif
    for (long row = 0; row < total_rows; ++row) {
        const int z = (int)(row / H), y = (int)(row % H);
        for (int x = 0; x < W; ++x) {
            float v[4];
            if (z >= used) {
                v[0] = p->background[0]; v[1] = p->background[1]; v[2] = p->background[2]; v[3] = p->background[3];
            } else {
                resize_pixel(&crops[z], &geoms[z], p->aspect_mode, p->src_type, p->yuv_standard, x, y, p->background, v);
            }
            apply_chain(p, v);
            store_pixel(p, n_planes, z, y, x, v);
        }
    }
    free(geoms);
    return 0;
}

/* Batched warp + chain + write: BatchRead<N, CONDITIONAL_WITH_DEFAULT> of Warping ops (cvGPUSpeedup.cuh:285-442). */
int oracle_warp(const cvgs_crop_t* images, const cvgs_warp_t* warps, int n_planes, int used, const cvgs_pipeline_t* p,
                int nthreads) {
    if (!images || !warps || !p || !p->out || n_planes <= 0 || used < 0 || CVGS_IS_YUV(p->src_type) || p->dst_width <= 0 ||
        p->dst_height <= 0 || p->n_ops < 0 || p->n_ops > CVGS_MAX_OPS)
        return 1;
    if (used > n_planes) used = n_planes;
    const int H = p->dst_height, W = p->dst_width;
    const long total_rows = (long)n_planes * H;
#ifdef _OPENMP
    if (nthreads <= 0) nthreads = omp_get_max_threads();
#pragma omp parallel for schedule(dynamic, 8) num_threads(nthreads)
#endif
    for (long row = 0; row < total_rows; ++row) {
        const int z = (int)(row / H), y = (int)(row % H);
        for (int x = 0; x < W; ++x) {
            float v[4];
            if (z >= used) {
                v[0] = p->background[0]; v[1] = p->background[1]; v[2] = p->background[2]; v[3] = p->background[3];
            } else {
                warp_pixel(&images[z], &warps[z], p->src_type, x, y, v);
            }
            apply_chain(p, v);
            store_pixel(p, n_planes, z, y, x, v);
        }
    }
    return 0;
}

/* ---------------------------------------------------------------------------------------------
 * CircularTensor state machine: fk::CircularTensor::update (circular_tensor.cuh:111-146),
 * SequenceSelectorType (:26-35), computeCircularThreadIdx (memory_operations.cuh:388-399).
 * Host memory; planes are [batch][color][H][W] (Standard) or [color][batch][H][W] (Transposed).
 * ------------------------------------------------------------------------------------------- */
typedef struct oracle_ct {
    int w, h, cp, ec, batch, order, mode, next;  /* cp colour planes of ec-channel elements */
    float* pub;  /* `this` tensor  */
    float* tmp;  /* m_tempTensor   */
} oracle_ct_t;

/* cp in {1, 3, 4} planes of floats (ec = 1), or one plane of packed ec = 3 / 4 float pixels (TensorWrite; the
 * reference's CircularTensor<CV_8UC4, CV_32FC4, 1, ...>, tests/batchread/test_circularbatchread_x_write3D.cu:400-460). */
oracle_ct_t* oracle_ct_create_ex(int w, int h, int cp, int ec, int batch, int order, int mode) {
    oracle_ct_t* t = (oracle_ct_t*)calloc(1, sizeof *t);
    t->w = w; t->h = h; t->cp = cp; t->ec = ec; t->batch = batch; t->order = order; t->mode = mode;
    const size_t n = (size_t)w * h * cp * ec * batch;
    t->pub = (float*)calloc(n, sizeof(float));
    t->tmp = (float*)calloc(n, sizeof(float));
    return t;
}
oracle_ct_t* oracle_ct_create(int w, int h, int cp, int batch, int order, int mode) {
    return oracle_ct_create_ex(w, h, cp, 1, batch, order, mode);
}
void oracle_ct_destroy(oracle_ct_t* t) { if (t) { free(t->pub); free(t->tmp); free(t); } }
float* oracle_ct_data(oracle_ct_t* t) { return t->pub; }
/* Lets a test preload the public tensor like the reference tests do (setTo(10.0f), :188-195). */
float* oracle_ct_temp(oracle_ct_t* t) { return t->tmp; }

static float* ct_plane(const oracle_ct_t* t, float* base, int z, int c) {
    const size_t px = (size_t)t->w * t->h * t->ec;
    return t->mode == CVGS_CT_STANDARD ? base + ((size_t)z * t->cp + c) * px
                                       : base + ((size_t)c * t->batch + z) * px;
}

int oracle_ct_update(oracle_ct_t* t, const cvgs_crop_t* frame, const cvgs_pipeline_t* p_in, int nthreads) {
    if (!t || !frame || !p_in || p_in->dst_width != t->w || p_in->dst_height != t->h) return 1;
    const size_t px = (size_t)t->w * t->h * t->ec;  /* floats per colour plane */
    float* fresh = (float*)malloc(px * t->cp * sizeof(float));
    cvgs_pipeline_t p = *p_in;
    /* the chain must end with cp * ec channels; packed elements are written pixel-interleaved */
    p.out = fresh; p.out_layout = t->ec > 1 ? CVGS_OUT_NHWC : CVGS_OUT_NCHW; p.out_plane_stride = 0; p.dst_type = 0;
    const int rc = oracle_preproc(frame, 1, 1, &p, nthreads);
    if (rc) { free(fresh); return rc; }
    const int B = t->batch, first = t->next;
    const int upd = t->order == CVGS_CT_NEWEST_FIRST ? 0 : B - 1;
    for (int c = 0; c < t->cp; ++c) {
        /* update sequence: MidWrite CircularTensorWrite<Ascendent> -> temp[(upd + first) mod B],
         * then the user's write -> this[upd]. */
        int zt = upd + first; if (zt >= B) zt -= B;
        memcpy(ct_plane(t, t->tmp, zt, c), fresh + c * px, px * sizeof(float));
        memcpy(ct_plane(t, t->pub, upd, c), fresh + c * px, px * sizeof(float));
        /* copy sequence for every other plane z. */
        for (int z = 0; z < B; ++z) {
            if (z == upd) continue;
            int zs;
            if (t->order == CVGS_CT_NEWEST_FIRST) { zs = first - z; if (zs < 0) zs += B; }
            else { zs = z + first; if (zs >= B) zs -= B; }
            memcpy(ct_plane(t, t->pub, z, c), ct_plane(t, t->tmp, zs, c), px * sizeof(float));
        }
    }
    t->next = (first + 1) % B;
    free(fresh);
    return 0;
}

int oracle_has_fma(void) { return __builtin_cpu_supports("fma") ? 1 : 0; }
int oracle_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
