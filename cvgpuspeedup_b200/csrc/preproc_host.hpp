// preproc_host.hpp -- host-side logic of the fused batch pipeline: argument validation,
// resize geometry, op-chain normalisation.  Pure C++ (no CUDA calls) so it is shared by the
// batch path and the CircularTensor path.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <string>

#include "../../include/cvgs_b200.h"
#include "cvgs_device.cuh"
#include "cvgs_runtime.hpp"

namespace cvgs {


// Host half of fk::Resize::build: scale factors and, for the aspect-ratio preserving modes, the
// band of the destination that receives the image (reference resize.cuh:100-161,191-216;
// rounding helper cxp::round, constexpr_cmath.cuh:37-48).
inline float round_half_away(float x) {
    if (x != x || std::isinf(x)) return x;
    return x > 0.f ? static_cast<float>(static_cast<int>(x + 0.5f)) : static_cast<float>(static_cast<int>(x - 0.5f));
}

inline void resize_geometry(int sw, int sh, int dw, int dh, int aspect, DevCrop& d) {
    int tw = dw, th = dh;
    if (aspect != CVGS_IGNORE_AR) {
        const float sf = dh / static_cast<float>(sh);
        int w_try = static_cast<int>(round_half_away(sf * sw));
        const bool even = aspect == CVGS_PRESERVE_AR_RN_EVEN;
        if (even) w_try -= w_try % 2;
        if (w_try > dw) {
            const float sf2 = dw / static_cast<float>(sw);
            int h_try = static_cast<int>(round_half_away(sf2 * sh));
            if (even) h_try -= h_try % 2;
            tw = dw;
            th = h_try;
        } else {
            tw = w_try;
            th = dh;
        }
    }
    const double cfx = static_cast<double>(tw) / static_cast<double>(sw);
    const double cfy = static_cast<double>(th) / static_cast<double>(sh);
    d.fx = static_cast<float>(1.0 / cfx);
    d.fy = static_cast<float>(1.0 / cfy);
    if (aspect == CVGS_IGNORE_AR) {
        d.bx1 = 0; d.by1 = 0; d.bx2 = dw - 1; d.by2 = dh - 1;
    } else {
        d.bx1 = aspect == CVGS_PRESERVE_AR_LEFT ? 0 : (dw - tw) / 2;
        d.by1 = (dh - th) / 2;
        d.bx2 = d.bx1 + tw - 1;
        d.by2 = d.by1 + th - 1;
    }
}

inline int channels_of(int src_type) { return (src_type == CVGS_8UC4 || src_type == CVGS_16UC4 || src_type == CVGS_16SC4) ? 4 : 3; }
// YCbCr -> RGB matrices of the reference (color_conversion.cuh:171-214), row-major, + luma offset
inline bool yuv_is_10bit(int src_type) { return src_type == CVGS_P010 || src_type == CVGS_P210 || src_type == CVGS_Y210; }
inline void yuv_constants(int standard, int src_type, float (&m)[12]) {
    static const float k[4][10] = {
        {1.164383562f, 0.f, 1.596026786f, 1.164383562f, -0.39176229f, -0.812967647f, 1.164383562f, 2.017232143f, 0.f, 16.f},
        {1.f, 0.f, 1.5748f, 1.f, -0.1873f, -0.4681f, 1.f, 1.8556f, 0.f, 0.f},
        {1.f, 0.f, 1.402f, 1.f, -0.34414f, -0.71414f, 1.f, 1.772f, 0.f, 0.f},
        {1.f, 0.f, 1.4746f, 1.f, -0.16455312684366f, -0.57135312684366f, 1.f, 1.8814f, 0.f, 0.f}};
    for (int i = 0; i < 10; ++i) m[i] = k[standard][i];
    // subCoefficients<p10bit> {64, 512}, floatShiftFactor<p10bit> 64 (color_conversion.cuh:106-111,189-192)
    const bool ten = yuv_is_10bit(src_type);
    if (ten) m[9] *= 4.f;
    m[10] = ten ? 512.f : 128.f;
    m[11] = ten ? 64.f : 1.f;
}
inline int pixel_bytes_of(int src_type) {
    switch (src_type) {
        case CVGS_NV12: case CVGS_NV21: return 1;  // luma plane
        case CVGS_P010: case CVGS_P210: return 2;
        case CVGS_Y210: return 4;                  // {Y0, U, Y1, V} words per pixel pair
        case CVGS_8UC3: return 3;
        case CVGS_8UC4: return 4;
        case CVGS_16UC3: case CVGS_16SC3: return 6;
        default: return 8;
    }
}

// Normalise the user's op list (header: enum cvgs_op_kind) into a DevProgram.
inline int build_program(const cvgs_pipeline_t& p, DevProgram& prog) {
    const int NC = channels_of(p.src_type);
    std::memset(&prog, 0, sizeof prog);
    if (p.n_ops < 0 || p.n_ops > CVGS_MAX_OPS) return fail(CVGS_ERR_INVALID_VALUE, "n_ops out of range");
    prog.round_u8 = p.interp_mode != CVGS_INTERP_ROUND_U8                        ? ROUND_NONE
                    : (p.src_type == CVGS_16UC3 || p.src_type == CVGS_16UC4)   ? ROUND_U16
                    : (p.src_type == CVGS_16SC3 || p.src_type == CVGS_16SC4)   ? ROUND_S16
                                                                               : ROUND_U8;
    int cur[4] = {0, 1, 2, 3};  // logical channel c currently lives in register cur[c] (registers = source channels)
    int ncur = NC;               // logical channels at this point of the chain
    const bool fuse = p.fp_contract == CVGS_FP_REFERENCE_FUSED;
    DevOp set{};                 // AddOpaqueAlpha, hoisted to the front (see below)
    bool have_set = false;
    DevOp body[CVGS_MAX_OPS];
    int n = 0;
    for (int i = 0; i < p.n_ops; ++i) {
        const cvgs_op_t& op = p.ops[i];
        if (op.kind == CVGS_OP_REORDER) {
            int nc[4] = {cur[0], cur[1], cur[2], cur[3]};
            bool seen[4] = {false, false, false, false};
            for (int c = 0; c < ncur; ++c) {
                if (op.perm[c] < 0 || op.perm[c] >= ncur || seen[op.perm[c]])
                    return fail(CVGS_ERR_INVALID_VALUE, "REORDER perm must be a permutation of the channels");
                seen[op.perm[c]] = true;
                nc[c] = cur[op.perm[c]];
            }
            std::memcpy(cur, nc, sizeof cur);
            continue;
        }
        if (op.kind == CVGS_OP_ADD_ALPHA) {
            if (ncur != 3) return fail(CVGS_ERR_INVALID_VALUE, "ADD_ALPHA needs a 3-channel pixel");
            if (have_set) return fail(CVGS_ERR_NOT_SUPPORTED, "one ADD_ALPHA per chain");
            const int r = 6 - cur[0] - cur[1] - cur[2];  // the register no live channel uses
            // The constant does not depend on anything computed before it, so the assignment is hoisted to the front
            // of the program and every earlier per-channel op gets the identity on that register (x*1, x+(-0), x/1
            // are exact): a MUL ... ADD pair around the conversion still contracts like in the reference's kernel.
            set.kind = DOP_SET;
            set.a[r] = op.v[0];
            set.b[r] = 1.f;
            have_set = true;
            for (int k = 0; k < n; ++k) {
                if ((body[k].kind & 0xff) == DOP_GRAY) continue;
                body[k].a[r] = body[k].kind == DOP_ADD ? -0.0f : 1.0f;
                body[k].b[r] = -0.0f;
            }
            cur[3] = r;
            ncur = 4;
            continue;
        }
        if (op.kind == CVGS_OP_DROP_ALPHA) {
            if (ncur != 4) return fail(CVGS_ERR_INVALID_VALUE, "DROP_ALPHA needs a 4-channel pixel");
            ncur = 3;
            continue;
        }
        if (n >= CVGS_MAX_OPS) return fail(CVGS_ERR_INVALID_VALUE, "too many operations");
        DevOp d{};
        if (op.kind == CVGS_OP_GRAY) {
            if (ncur < 3) return fail(CVGS_ERR_INVALID_VALUE, "GRAY needs a 3- or 4-channel pixel");
            if (op.perm[0] != 0 && op.perm[0] != 1) return fail(CVGS_ERR_INVALID_VALUE, "GRAY: perm[0] is 0 or 1");
            d.kind = DOP_GRAY | (cur[0] << 8) | (cur[1] << 12) | (cur[2] << 16) | (fuse ? 0 : 1 << 20) | (op.perm[0] << 21);
            body[n++] = d;
            cur[0] = 0;  // DOP_GRAY leaves its result in register 0
            ncur = 1;
            continue;
        }
        switch (op.kind) {
            case CVGS_OP_MUL: d.kind = DOP_MUL; break;
            case CVGS_OP_DIV: d.kind = DOP_DIV; break;
            case CVGS_OP_ADD:
            case CVGS_OP_SUB: d.kind = DOP_ADD; break;
            default: return fail(CVGS_ERR_INVALID_VALUE, "unknown op kind");
        }
        // registers that hold no live channel get the identity (their lanes are computed but never stored)
        for (int r = 0; r < 4; ++r) d.a[r] = d.kind == DOP_ADD ? -0.0f : 1.0f;
        for (int c = 0; c < ncur; ++c) d.a[cur[c]] = op.kind == CVGS_OP_SUB ? -op.v[c] : op.v[c];
        // nvcc contracts (x*m) +/- s into one FMA in the reference's inlined chain
        if (fuse && d.kind == DOP_ADD && n > 0 && body[n - 1].kind == DOP_MUL) {
            DevOp& m = body[n - 1];
            m.kind = DOP_FMA;
            for (int r = 0; r < 4; ++r) m.b[r] = d.a[r];
            continue;
        }
        body[n++] = d;
    }
    if (n + (have_set ? 1 : 0) > CVGS_MAX_OPS) return fail(CVGS_ERR_INVALID_VALUE, "too many operations");
    int k = 0;
    if (have_set) prog.ops[k++] = set;
    for (int i = 0; i < n; ++i) prog.ops[k++] = body[i];
    prog.n_ops = k;
    for (int r = 0; r < 4; ++r) prog.dst_chan[r] = -1;
    for (int c = 0; c < ncur; ++c) prog.dst_chan[cur[c]] = c;
    prog.nc_out = ncur;
    prog.nregs = (have_set && NC == 3) ? 4 : NC;
    prog.special = (have_set || ncur != NC) ? 1 : 0;
    for (int i = 0; i < k; ++i)
        if ((prog.ops[i].kind & 0xff) == DOP_GRAY) prog.special = 1;
    return CVGS_OK;
}

inline int validate_pipeline(const cvgs_pipeline_t* p) {
    if (!p) return fail(CVGS_ERR_INVALID_VALUE, "pipeline is NULL");
    if (p->src_type != CVGS_8UC3 && p->src_type != CVGS_16UC3 && p->src_type != CVGS_16SC3 && p->src_type != CVGS_8UC4 &&
        p->src_type != CVGS_16UC4 && p->src_type != CVGS_16SC4 && !CVGS_IS_YUV(p->src_type))
        return fail(CVGS_ERR_NOT_SUPPORTED, "sources must be CV_8U / CV_16U / CV_16S with 3 or 4 channels, or NV12 / NV21 / P010 / P210 / Y210 frames");
    if (CVGS_IS_YUV(p->src_type) && (p->yuv_standard < 0 || p->yuv_standard > 3))
        return fail(CVGS_ERR_INVALID_VALUE, "bad yuv_standard");
    if (p->dst_width <= 0 || p->dst_height <= 0 || p->dst_width > (1 << 20) || p->dst_height > (1 << 20))
        return fail(CVGS_ERR_INVALID_VALUE, "destination size out of range");
    if (p->aspect_mode < 0 || p->aspect_mode > 3) return fail(CVGS_ERR_INVALID_VALUE, "bad aspect_mode");
    if (p->interp_mode < 0 || p->interp_mode > 1) return fail(CVGS_ERR_INVALID_VALUE, "bad interp_mode");
    if (p->fp_contract < 0 || p->fp_contract > 1) return fail(CVGS_ERR_INVALID_VALUE, "bad fp_contract");
    if (p->out_layout < 0 || p->out_layout > 3) return fail(CVGS_ERR_INVALID_VALUE, "bad out_layout");
    if (p->out_plane_stride < 0) return fail(CVGS_ERR_INVALID_VALUE, "negative out_plane_stride");
    if (p->dst_type != 0 && p->dst_type != CVGS_32FC1 && p->dst_type != CVGS_32FC3 && p->dst_type != CVGS_32FC4 &&
        p->dst_type != CVGS_8UC3 && p->dst_type != CVGS_8UC4)
        return fail(CVGS_ERR_NOT_SUPPORTED, "dst_type must be CV_32FC1 / CV_32FC3 / CV_32FC4 (or 0) or CV_8UC3 / CV_8UC4");
    if (p->u8_cast != 0 && (p->u8_cast != 1 || (p->dst_type != CVGS_8UC3 && p->dst_type != CVGS_8UC4)))
        return fail(CVGS_ERR_INVALID_VALUE, "u8_cast is 0 or 1 and applies to 8-bit output");
    const bool u8_out = p->dst_type == CVGS_8UC3 || p->dst_type == CVGS_8UC4;
    if (u8_out && p->out_layout != CVGS_OUT_NHWC)
        return fail(CVGS_ERR_INVALID_VALUE, "8-bit output is packed: out_layout must be CVGS_OUT_NHWC");
    if (u8_out && p->out_row_pitch != 0 && p->out_row_pitch < (p->dst_type == CVGS_8UC3 ? 3LL : 4LL) * p->dst_width)
        return fail(CVGS_ERR_INVALID_VALUE, "out_row_pitch smaller than a row");
    if (!u8_out && p->out_row_pitch != 0 && p->out_layout != CVGS_OUT_NHWC)
        return fail(CVGS_ERR_INVALID_VALUE, "out_row_pitch applies to packed output (CVGS_OUT_NHWC)");
    if (p->out_row_pitch < 0) return fail(CVGS_ERR_INVALID_VALUE, "negative out_row_pitch");
    return CVGS_OK;
}

// Wide stores: planar float rows written 16 bytes at a time, or (8-bit output) quads of pixels as whole 32-bit words.
inline int out_vec4(const OutDesc& o, int dst_width, bool planes, const void* out) {
    const uintptr_t a = reinterpret_cast<uintptr_t>(out);
    if (o.u8) return (a % 4) == 0 && (o.row_pitch % 4) == 0 && (o.z_stride % 4) == 0;
    return !planes && o.px_stride == 1 && (dst_width % 4) == 0 && (a % 16) == 0 && (o.z_stride % 4) == 0 && (o.c_stride % 4) == 0;
}

// Fill the launch parameters that do not depend on the crops.
inline int build_params(const cvgs_pipeline_t& p, int n_planes, int used, float* out, PreprocParams& P) {
    std::memset(&P, 0, sizeof P);
    P.n_planes = n_planes;
    P.used = used;
    P.W = p.dst_width;
    P.H = p.dst_height;
    P.band_test = p.aspect_mode != CVGS_IGNORE_AR;
    P.src_type = p.src_type;
    if (CVGS_IS_YUV(p.src_type)) yuv_constants(p.yuv_standard, p.src_type, P.yuv);
    P.nc = channels_of(p.src_type);
    for (int c = 0; c < P.nc; ++c) P.bg[c] = p.background[c];
    if (int rc = build_program(p, P.prog)) return rc;
    const int NC = P.prog.nc_out;  // channels of the OUTPUT pixel
    const bool u8_out = p.dst_type == CVGS_8UC3 || p.dst_type == CVGS_8UC4;
    if (P.prog.special && p.out_layout == CVGS_OUT_PLANES)
        return fail(CVGS_ERR_NOT_SUPPORTED, "conversions that change the channel count write tensors or packed images, not per-plane images");
    if (u8_out && NC != (p.dst_type == CVGS_8UC3 ? 3 : 4))
        return fail(CVGS_ERR_INVALID_VALUE, "dst_type does not match the channels the chain produces");
    if (p.dst_type == CVGS_32FC1 || p.dst_type == CVGS_32FC3 || p.dst_type == CVGS_32FC4) {
        const int want = p.dst_type == CVGS_32FC1 ? 1 : (p.dst_type == CVGS_32FC3 ? 3 : 4);
        if (want != NC) return fail(CVGS_ERR_INVALID_VALUE, "dst_type does not match the channels the chain produces");
    }
    const long long plane = static_cast<long long>(p.dst_width) * p.dst_height;
    OutDesc& o = P.out;
    o.base = out;
    o.planes = nullptr;
    switch (p.out_layout) {
        case CVGS_OUT_PLANES:  // the launch path uploads the plane table and sets o.planes
            o.base = nullptr;
            o.z_stride = 0;
            o.c_stride = 0;
            o.px_stride = 1;
            break;
        case CVGS_OUT_NCHW:
            o.z_stride = p.out_plane_stride ? p.out_plane_stride : NC * plane;
            o.c_stride = plane;
            o.px_stride = 1;
            break;
        case CVGS_OUT_CNHW:
            o.z_stride = p.out_plane_stride ? p.out_plane_stride : plane;
            o.c_stride = o.z_stride * n_planes;
            o.px_stride = 1;
            break;
        default:
            o.z_stride = p.out_plane_stride ? p.out_plane_stride : NC * plane;
            o.c_stride = 1;
            o.px_stride = NC;
    }
    o.u8 = u8_out ? (p.u8_cast ? 2 : 1) : 0;
    o.row_pitch = 0;
    o.row_stride = static_cast<long long>(p.dst_width) * o.px_stride;
    if (o.u8) {  // strides in bytes
        o.row_pitch = p.out_row_pitch ? p.out_row_pitch : static_cast<long long>(NC) * p.dst_width;
        o.z_stride = p.out_plane_stride ? p.out_plane_stride : o.row_pitch * p.dst_height;
    } else if (p.out_row_pitch) {  // packed float image with padded rows (a GpuMat from cudaMallocPitch)
        if ((p.out_row_pitch & 3) || p.out_row_pitch < 4LL * NC * p.dst_width)
            return fail(CVGS_ERR_INVALID_VALUE, "out_row_pitch must be a multiple of 4 and hold a row");
        o.row_stride = p.out_row_pitch / 4;
        o.z_stride = p.out_plane_stride ? p.out_plane_stride : o.row_stride * p.dst_height;
    }
    for (int r = 0; r < 4; ++r) o.c_off[r] = P.prog.dst_chan[r] < 0 ? kNoStore : P.prog.dst_chan[r] * o.c_stride;
    o.vec4 = out_vec4(o, p.dst_width, p.out_layout == CVGS_OUT_PLANES, out);
    return CVGS_OK;
}

inline int fill_crop(const cvgs_crop_t& c, const cvgs_pipeline_t& p, int idx, DevCrop& d) {
    if (!c.data) return fail(CVGS_ERR_INVALID_VALUE, "crop " + std::to_string(idx) + ": data is NULL");
    if (c.width <= 0 || c.height <= 0 || c.width > (1 << 22) || c.height > (1 << 22))
        return fail(CVGS_ERR_INVALID_VALUE, "crop " + std::to_string(idx) + ": size out of range");
    const int px_bytes = pixel_bytes_of(p.src_type);
    if (c.pitch < px_bytes * c.width && c.height > 1)
        return fail(CVGS_ERR_INVALID_VALUE, "crop " + std::to_string(idx) + ": pitch smaller than a row");
    const int align = p.src_type == CVGS_Y210 ? 8 : ((p.src_type == CVGS_P010 || p.src_type == CVGS_P210) ? 4 : (px_bytes >= 6 ? 2 : 1));
    if ((reinterpret_cast<uintptr_t>(c.data) | static_cast<uintptr_t>(c.pitch)) & (align - 1))
        return fail(CVGS_ERR_INVALID_VALUE, "crop " + std::to_string(idx) + ": 16-bit pixels must be aligned to their vector type");
    d.data = static_cast<const uint8_t*>(c.data);
    d.w = c.width;
    d.h = c.height;
    d.pitch = c.pitch;
    d.pad = 0;
    resize_geometry(c.width, c.height, p.dst_width, p.dst_height, p.aspect_mode, d);
    return CVGS_OK;
}

}  // namespace cvgs
