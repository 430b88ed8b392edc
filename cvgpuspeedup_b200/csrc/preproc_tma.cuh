// preproc_tma.cuh -- persistent, warp-specialised, TMA-staged kernel of the fused batch path (sm_100a).
//
// Same function as preproc_direct_kernel (preproc.cu) and as the reference instantiation of
// fk::launchTransformDPP_Kernel it replaces (reference fkl/include/fused_kernel/core/execution_model/
// data_parallel_patterns.cuh:157-197; BatchRead batch_operations.cuh:222-229; Resize resize.cuh:70-82,178-189;
// Interpolate interpolation.cuh:57-92; TensorSplit memory_operations.cuh:168-188), organised around the
// issue-slot budget of the path:
//
//   * a CTA walks a contiguous range of tiles; a tile = TR output rows x TW output columns of one crop;
//   * the producer warp stages, per output row of the tile, the two source rows it taps with ONE
//     cp.async.bulk.tensor.2d (box = 2 rows x the tile's source span; byte-exact start coordinate through a
//     per-crop tensor map of 8-byte elements), completion on a per-stage mbarrier, kStages deep;
//   * consumer threads own one quad (4 adjacent output columns) for as long as the CTA stays on the same
//     crop column band, so the horizontal taps are computed once and only the vertical ones per row;
//   * u8 -> f32 is one PRMT: byte b placed at bits 16..23 of a float word is exactly b * 2^-133 (exponent
//     field 0 or 1, both scale 2^-149); the vertical weights carry 2^100 and the first op of the chain the
//     remaining 2^33 (kPreScale).  Scaling by powers of two commutes with every rounding below (no value
//     leaves the normal range), so results are bit-identical to the unscaled sequence.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>

#include "../../include/cvgs_b200.h"
#include "cvgs_device.cuh"
#include "cvgs_runtime.hpp"

namespace cvgs {

constexpr int kTmaParamCrops = 64;      // crops (descriptor + tensor map) that ride in the kernel parameters
constexpr int kConsumerWarps = 4;
constexpr int kConsumerThreads = kConsumerWarps * 32;
constexpr int kTmaThreads = kConsumerThreads + 32;  // + producer warp (the last one)
constexpr int kStages = 4;
constexpr int kStageBytesMax = 26 * 1024;
constexpr int kStagePad = 128;          // bytes in front of / behind the stage ring (w[-1] / w[+1] over-reads)
constexpr int kMaxBoxBytes = 2048;      // 256 elements x 8 bytes
constexpr float kWeightScale = 1.2676506002282294e30f;  // 2^100
constexpr float kPreScale = 8589934592.0f;              // 2^33  = 2^133 / 2^100

struct TmaGeom {
    int32_t TW;              // tile width in output pixels (power of two, 8..256)
    int32_t q_log2;          // log2(quads per tile row)
    int32_t groups;          // rows processed concurrently by the consumer threads = 128 >> q_log2
    int32_t TR;              // output rows per tile
    int32_t tiles_x, tiles_y, tiles_per_crop, total_tiles;
    int32_t stage_bytes;     // TR * 2 * max row bytes
    int32_t stages;          // depth of the stage ring actually used (1..kStages)
    int32_t grid;            // CTAs; CTA b walks tiles [b*tiles_base + min(b, tiles_rem), ...)
    int32_t tiles_base, tiles_rem;
    int32_t explicit_prescale;  // 1: consumers multiply by 2^33 themselves (no op to fold it into)
};

struct TmaParams {
    PreprocParams P;         // P.prog = unscaled chain (background values), P.crops = device table or nullptr
    DevProgram prog_img;     // chain for interpolated values (2^33 folded into its first op)
    float div_rcp[3];        // CH_FMA_DIV: RN(1/d) per source channel (exact-division fast path, see div_by_const)
    TmaGeom G;
    const CUtensorMap* maps; // device table (nullptr when the maps ride in the kernel parameters)
};

struct alignas(64) TmaParamTable {
    CUtensorMap m[kTmaParamCrops];
    DevCrop c[kTmaParamCrops];
};
struct TmaNoTable {
    int32_t unused;
};

// DevCrop::pad of a TMA launch: bits 0..15 = smem row bytes of this crop's box, bits 16..19 = data & 15
__host__ __device__ __forceinline__ int32_t crop_row_bytes(const DevCrop& c) { return c.pad & 0xFFFF; }
__host__ __device__ __forceinline__ int32_t crop_misalign(const DevCrop& c) { return (c.pad >> 16) & 15; }

// ------------------------------------------------------------------------------------------------
// PTX helpers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}" ::"r"(bar),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int32_t c0, int32_t c1, uint32_t bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
        "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(bar)
        : "memory");
}
__device__ __forceinline__ uint32_t lds32(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}

// byte k of w as the float b * 2^-133 (see file header)
__device__ __forceinline__ float u8_scaled(uint32_t w, uint32_t k) {
    return __uint_as_float(__byte_perm(w, 0u, 0x4044u | (k << 8)));
}

template <typename Table>
__device__ __forceinline__ const DevCrop& tma_crop_of(const TmaParams& K, const Table& T, int z);
template <>
__device__ __forceinline__ const DevCrop& tma_crop_of<TmaParamTable>(const TmaParams&, const TmaParamTable& T, int z) {
    return T.c[z];
}
template <>
__device__ __forceinline__ const DevCrop& tma_crop_of<TmaNoTable>(const TmaParams& K, const TmaNoTable&, int z) {
    return K.P.crops[z];
}
template <typename Table>
__device__ __forceinline__ const CUtensorMap* tma_map_of(const TmaParams& K, const Table& T, int z);
template <>
__device__ __forceinline__ const CUtensorMap* tma_map_of<TmaParamTable>(const TmaParams&, const TmaParamTable& T, int z) {
    return &T.m[z];
}
template <>
__device__ __forceinline__ const CUtensorMap* tma_map_of<TmaNoTable>(const TmaParams& K, const TmaNoTable&, int z) {
    return K.maps + z;
}

// Where the staged span of a (crop, column band) starts: first in-band output column of the band, its left tap,
// and the 8-byte element coordinate of the box.  Producer and consumers must agree, so both call this.
struct BandOrigin {
    int32_t xa;       // first output column of the band that receives image data (may exceed the band: empty)
    int32_t xe;       // last such column
    int32_t c0;       // box start coordinate (8-byte elements) in the crop's tensor map
    int32_t origin;   // crop-row byte that smem byte 0 of a staged row corresponds to (= 8*c0 - misalign)
};
__device__ __forceinline__ BandOrigin band_origin(const PreprocParams& P, const TmaGeom& G, const DevCrop& C, int txi) {
    BandOrigin b;
    const int tx0 = txi * G.TW;
    b.xa = max(tx0, C.bx1);
    b.xe = min(min(tx0 + G.TW, P.W) - 1, C.bx2);
    const AxisTap t = axis_tap(b.xa - C.bx1, C.fx);
    const int mis = crop_misalign(C);
    b.c0 = ((mis + 3 * t.i1) >> 4) << 1;  // the box must start on a 16-byte boundary of global memory
    b.origin = 8 * b.c0 - mis;
    return b;
}

// ------------------------------------------------------------------------------------------------
// kernel
// ------------------------------------------------------------------------------------------------
enum ChainKind : int { CH_GENERIC = 0, CH_FMA_DIV = 1 };

// Correctly rounded x / d for a launch constant d with r = RN(1/d) and nd = -d precomputed (Markstein): the first
// correction makes the quotient faithful, the second rounds it correctly; both residuals are exact FMAs.
// Equal to __fdiv_rn(x, d) bit for bit whenever 2^-60 <= |x| <= 2^60 and 2^-30 <= |d| <= 2^30 (no intermediate
// leaves the normal range); the caller checks the range of x per quad and uses __fdiv_rn otherwise.
// tests/test_division_gpu.py sweeps all 2^32 values of x against __fdiv_rn for a set of divisors.
constexpr float kDivSafeMin = 8.6736174e-19f;   // 2^-60
constexpr float kDivSafeMax = 1.1529215e18f;    // 2^60
__device__ __forceinline__ float div_by_const(float x, float r, float nd) {
    const float q0 = __fmul_rn(x, r);
    const float e0 = __fmaf_rn(q0, nd, x);
    const float q1 = __fmaf_rn(e0, r, q0);
    const float e1 = __fmaf_rn(q1, nd, x);
    return __fmaf_rn(e1, r, q1);
}

// Tile walk of one CTA: contiguous range, decoded once and then advanced incrementally (no divisions per tile).
struct TileCursor {
    int z, txi, tyi, left;
    __device__ __forceinline__ void init(const TmaGeom& G) {
        const int b = blockIdx.x;
        const int t0 = b * G.tiles_base + min(b, G.tiles_rem);
        left = G.tiles_base + (b < G.tiles_rem ? 1 : 0);
        z = t0 / G.tiles_per_crop;
        const int rem = t0 - z * G.tiles_per_crop;
        txi = rem / G.tiles_y;
        tyi = rem - txi * G.tiles_y;
    }
    __device__ __forceinline__ void next(const TmaGeom& G) {
        --left;
        if (++tyi == G.tiles_y) {
            tyi = 0;
            if (++txi == G.tiles_x) {
                txi = 0;
                ++z;
            }
        }
    }
};

// One output pixel: taps from the two staged rows -> 3 interpolated channels (scaled by 2^-33).
__device__ __forceinline__ void gather_px(uint32_t aA, uint32_t aB, int shl, int shr, bool edge, float wx0, float wx1,
                                          float wy0, float wy1, float (&v)[3]) {
    const uint32_t am = lds32(aA - 4), a0 = lds32(aA), a1 = lds32(aA + 4);
    const uint32_t bm = lds32(aB - 4), b0 = lds32(aB), b1 = lds32(aB + 4);
    const uint32_t al = __funnelshift_rc(am, a0, shl);  // left pixel in bytes 0..2 (clamped shift: 32 = a0 itself)
    const uint32_t bl = __funnelshift_rc(bm, b0, shl);
    uint32_t ar = __funnelshift_r(a0, a1, shr);         // right pixel in bytes 0..2
    uint32_t br = __funnelshift_r(b0, b1, shr);
    if (edge) {  // x2_read == x1 (interpolation.cuh:72): the right tap is the left pixel again
        ar = al;
        br = bl;
    }
    const float w00 = __fmul_rn(wx0, wy0), w10 = __fmul_rn(wx1, wy0);
    const float w01 = __fmul_rn(wx0, wy1), w11 = __fmul_rn(wx1, wy1);
#pragma unroll
    for (int c = 0; c < 3; ++c)
        v[c] = bilerp(u8_scaled(al, c), u8_scaled(ar, c), u8_scaled(bl, c), u8_scaled(br, c), w00, w10, w01, w11);
}

// GEN = false: the common geometry -- IGNORE_AR, every plane used, planar output with 16-byte aligned rows.
// GEN = true : aspect-ratio bands, unused planes, packed / unaligned outputs.
template <typename Table, int CHAIN, bool GEN>
__global__ void __launch_bounds__(kTmaThreads, GEN ? 2 : 4)
preproc_tma_kernel(const __grid_constant__ TmaParams K, const __grid_constant__ Table T) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ uint64_t bar_full[kStages];
    __shared__ uint64_t bar_empty[kStages];

    const PreprocParams& P = K.P;
    const TmaGeom& G = K.G;
    const int tid = threadIdx.x;
    const int warp = tid >> 5;
    const int lane = tid & 31;
    const int nstages = G.stages;

    // stage ring, 128-byte aligned, with kStagePad bytes of slack on both sides
    const uint32_t ring = ((smem_u32(smem_raw) + 127u) & ~127u) + kStagePad;

    if (tid == 0) {
        for (int s = 0; s < nstages; ++s) {
            mbar_init(smem_u32(&bar_full[s]), 1);
            mbar_init(smem_u32(&bar_empty[s]), kConsumerWarps);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    TileCursor tc;
    tc.init(G);

    if (warp == kConsumerWarps) {
        // ===================================== producer warp =====================================
        int stage = 0;
        uint32_t phase = 0;
        for (; tc.left > 0; tc.next(G)) {
            mbar_wait(smem_u32(&bar_empty[stage]), phase ^ 1u);
            const uint32_t full = smem_u32(&bar_full[stage]);
            bool issue = false;
            int i1 = 0, c0 = 0, rb = 0;
            if (!GEN || tc.z < P.used) {
                const DevCrop& C = tma_crop_of<Table>(K, T, tc.z);
                const BandOrigin b = band_origin(P, G, C, tc.txi);
                const int y = tc.tyi * G.TR + lane;
                rb = crop_row_bytes(C);
                if (lane < G.TR && y < P.H && y >= C.by1 && y <= C.by2 && b.xa <= b.xe) {
                    issue = true;
                    i1 = axis_tap(y - C.by1, C.fy).i1;
                    c0 = b.c0;
                }
            }
            const unsigned m = __ballot_sync(0xffffffffu, issue);
            if (lane == 0) mbar_arrive_expect_tx(full, (uint32_t)(__popc(m) * 2 * rb));
            __syncwarp();
            if (issue) {
                const CUtensorMap* map = tma_map_of<Table>(K, T, tc.z);
                if (K.maps)  // table written by a host copy into reused ring memory: acquire it for the TMA proxy
                    asm volatile("fence.proxy.tensormap::generic.acquire.gpu [%0], 128;" ::"l"(reinterpret_cast<uint64_t>(map))
                                 : "memory");
                tma_load_2d(ring + stage * G.stage_bytes + lane * 2 * rb, map, c0, i1, full);
            }
            if (++stage == nstages) {
                stage = 0;
                phase ^= 1u;
            }
        }
        return;
    }

    // ========================================= consumers =========================================
    const int qx = tid & ((1 << G.q_log2) - 1);
    const int g = tid >> G.q_log2;
    const int W = P.W, H = P.H;

    // chain constants of the specialised shape v = fma(v, ca, cb) / cd  (source-channel order)
    float ca[3], cb[3], cd[3], cr[3];
    if (CHAIN == CH_FMA_DIV) {
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            ca[c] = K.prog_img.ops[0].a[c];
            cb[c] = K.prog_img.ops[0].b[c];
            cd[c] = K.prog_img.ops[1].a[c];
            cr[c] = K.div_rcp[c];
        }
    }
    // chain(background): value of planes z >= used and of pixels outside the aspect-ratio band
    float vb[1][3] = {{P.bg[0], P.bg[1], P.bg[2]}};
    if (GEN) apply_program<1>(P.prog, vb);

    // planar fast path: plane offsets (floats) of the three source channels
    const long long oc0 = (long long)P.prog.dst_chan[0] * P.out.c_stride;
    const long long oc1 = (long long)P.prog.dst_chan[1] * P.out.c_stride;
    const long long oc2 = (long long)P.prog.dst_chan[2] * P.out.c_stride;

    // per-thread horizontal state, valid while (z, txi) is unchanged
    int cur_band = -1;
    int32_t off0 = 0, off1 = 0, off2 = 0, off3 = 0, shl0 = 0, shl1 = 0, shl2 = 0, shl3 = 0, shr0 = 0, shr1 = 0, shr2 = 0,
            shr3 = 0;
    float wxa0 = 0, wxa1 = 0, wxa2 = 0, wxa3 = 0, wxb0 = 0, wxb1 = 0, wxb2 = 0, wxb3 = 0;
    bool e0 = false, e1 = false, e2 = false, e3 = false;
    bool n0 = false, n1 = false, n2 = false, n3 = false;  // GEN: pixel p receives image data
    bool band_ok = false;
    int nvalid = 0, x0 = 0;
    // crop fields used per row
    float c_fy = 0.f;
    int c_by1 = 0, c_by2 = 0, c_hm1 = 0, c_rb = 0;

    int stage = 0;
    uint32_t phase = 0;
    for (; tc.left > 0; tc.next(G)) {
        const int z = tc.z;
        const bool active = !GEN || z < P.used;
        const int band = z * G.tiles_x + tc.txi;
        if (band != cur_band) {
            cur_band = band;
            x0 = tc.txi * G.TW + 4 * qx;
            nvalid = min(4, W - x0);  // <= 0: this thread has no column in the tile
            band_ok = false;
            if (active) {
                const DevCrop& C = tma_crop_of<Table>(K, T, z);
                const BandOrigin b = band_origin(P, G, C, tc.txi);
                band_ok = b.xa <= b.xe;
                c_fy = C.fy;
                c_by1 = C.by1;
                c_by2 = C.by2;
                c_hm1 = C.h - 1;
                c_rb = crop_row_bytes(C);
                const float fx = C.fx;
                const int bx1 = C.bx1, wm1 = C.w - 1, origin = b.origin;
#define CVGS_XSETUP(p, OFF, SHL, SHR, WA, WB, E, N)                                     \
    {                                                                                   \
        const int x = x0 + p;                                                           \
        N = p < nvalid && x >= b.xa && x <= b.xe;                                       \
        const AxisTap t_ = axis_tap((N ? x : b.xa) - bx1, fx);                          \
        WA = t_.w0;                                                                     \
        WB = t_.w1;                                                                     \
        E = t_.i1 + 1 > wm1;                                                            \
        const int o = 3 * t_.i1 - origin;                                               \
        OFF = ((o + 3) >> 2) * 4;                                                       \
        SHL = (o & 3) ? (o & 3) * 8 : 32;                                               \
        SHR = ((o + 3) & 3) * 8;                                                        \
    }
                CVGS_XSETUP(0, off0, shl0, shr0, wxa0, wxb0, e0, n0)
                CVGS_XSETUP(1, off1, shl1, shr1, wxa1, wxb1, e1, n1)
                CVGS_XSETUP(2, off2, shl2, shr2, wxa2, wxb2, e2, n2)
                CVGS_XSETUP(3, off3, shl3, shr3, wxa3, wxb3, e3, n3)
#undef CVGS_XSETUP
            }
        }

        const int ybase = tc.tyi * G.TR;
        // planar fast path: the three channel planes of this thread's quad at row ybase
        float* const tp = P.out.base + ((long long)z * P.out.z_stride + (long long)ybase * W + x0);
        float* const tp0 = tp + oc0;
        float* const tp1 = tp + oc1;
        float* const tp2 = tp + oc2;

        mbar_wait(smem_u32(&bar_full[stage]), phase);
        const uint32_t sbase = ring + stage * G.stage_bytes;

        if (nvalid > 0) {
            for (int r = g; r < G.TR; r += G.groups) {
                const int y = ybase + r;
                if (y >= H) break;
                float v[4][3];
                const bool row_in = GEN ? (band_ok && y >= c_by1 && y <= c_by2) : true;
                if (row_in) {
                    const AxisTap ty_ = axis_tap(y - c_by1, c_fy);
                    const uint32_t rowA = sbase + r * 2 * c_rb;
                    const uint32_t rowB = rowA + ((ty_.i1 + 1 > c_hm1) ? 0 : c_rb);
                    const float wy0 = __fmul_rn(ty_.w0, kWeightScale);
                    const float wy1 = __fmul_rn(ty_.w1, kWeightScale);
                    gather_px(rowA + off0, rowB + off0, shl0, shr0, e0, wxa0, wxb0, wy0, wy1, v[0]);
                    gather_px(rowA + off1, rowB + off1, shl1, shr1, e1, wxa1, wxb1, wy0, wy1, v[1]);
                    gather_px(rowA + off2, rowB + off2, shl2, shr2, e2, wxa2, wxb2, wy0, wy1, v[2]);
                    gather_px(rowA + off3, rowB + off3, shl3, shr3, e3, wxa3, wxb3, wy0, wy1, v[3]);
                    if (CHAIN == CH_FMA_DIV) {
                        float amin = kDivSafeMax;  // |v| <= 2^60 is guaranteed by the host (bounded constants)
#pragma unroll
                        for (int p = 0; p < 4; ++p)
#pragma unroll
                            for (int c = 0; c < 3; ++c) {
                                v[p][c] = __fmaf_rn(v[p][c], ca[c], cb[c]);
                                amin = fminf(amin, fabsf(v[p][c]));
                            }
                        if (amin >= kDivSafeMin) {
#pragma unroll
                            for (int p = 0; p < 4; ++p)
#pragma unroll
                                for (int c = 0; c < 3; ++c) v[p][c] = div_by_const(v[p][c], cr[c], -cd[c]);
                        } else {  // zeros / denormal-range values: the IEEE routine
#pragma unroll
                            for (int p = 0; p < 4; ++p)
#pragma unroll
                                for (int c = 0; c < 3; ++c) v[p][c] = __fdiv_rn(v[p][c], cd[c]);
                        }
                    } else {
                        if (G.explicit_prescale) {
#pragma unroll
                            for (int p = 0; p < 4; ++p)
#pragma unroll
                                for (int c = 0; c < 3; ++c) v[p][c] = __fmul_rn(v[p][c], kPreScale);
                        }
                        apply_program<4>(K.prog_img, v);
                    }
                    if (GEN) {
#pragma unroll
                        for (int c = 0; c < 3; ++c) {
                            if (!n0) v[0][c] = vb[0][c];
                            if (!n1) v[1][c] = vb[0][c];
                            if (!n2) v[2][c] = vb[0][c];
                            if (!n3) v[3][c] = vb[0][c];
                        }
                    }
                } else {
#pragma unroll
                    for (int p = 0; p < 4; ++p)
#pragma unroll
                        for (int c = 0; c < 3; ++c) v[p][c] = vb[0][c];
                }
                if (GEN) {
                    store_pixels<4>(P, z, y, x0, nvalid, v);
                } else {
                    const int ro = r * W;
                    st_cs_f32x4(tp0 + ro, v[0][0], v[1][0], v[2][0], v[3][0]);
                    st_cs_f32x4(tp1 + ro, v[0][1], v[1][1], v[2][1], v[3][1]);
                    st_cs_f32x4(tp2 + ro, v[0][2], v[1][2], v[2][2], v[3][2]);
                }
            }
        }

        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&bar_empty[stage]));
        if (++stage == nstages) {
            stage = 0;
            phase ^= 1u;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
using EncodeTiledFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn encode_tiled_fn() {
    static EncodeTiledFn fn = [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess)
            p = nullptr;
        return reinterpret_cast<EncodeTiledFn>(p);
    }();
    return fn;
}

// Source bytes a staged row of a TW-wide band can span for scale factor fx (+ alignment slack), rounded to
// the 64 bytes that keep every 2-row box 128-byte aligned in shared memory.
inline int band_row_bytes(int TW, float fx) {
    // taps of TW columns: floor((TW-1)*fx) + 2 pixels, +1 for the float rounding of the two products
    const double px = std::ceil(static_cast<double>(TW - 1) * static_cast<double>(fx)) + 3.0;
    const long long bytes = static_cast<long long>(px) * 3 + 15 /*16-byte aligned start*/ + 4 /*w[+1] word*/;
    return static_cast<int>((bytes + 63) / 64 * 64);
}

// Can this launch take the TMA kernel, and with which tiling?  crops = host copies of the DevCrops.
inline bool tma_plan(const PreprocParams& P, const DevCrop* crops, int used, int n_planes, int sm_count, TmaGeom& G) {
    if (!encode_tiled_fn()) return false;
    float fx_max = 0.f;
    for (int i = 0; i < used; ++i) {
        const DevCrop& c = crops[i];
        if (c.h > 1 && c.pitch % 16 != 0) return false;  // TMA: row stride must be a multiple of 16 bytes
        if (!(c.fx > 0.f) || !(c.fy > 0.f) || !std::isfinite(c.fx) || !std::isfinite(c.fy)) return false;
        fx_max = std::max(fx_max, c.fx);
    }
    int TW = 256;
    while (TW > 8 && (TW / 2 >= P.W || band_row_bytes(std::min(TW, P.W), fx_max) > kMaxBoxBytes)) TW /= 2;
    if (used > 0 && band_row_bytes(std::min(TW, P.W), fx_max) > kMaxBoxBytes) return false;  // extreme down-scale: direct kernel
    const int rb_max = used > 0 ? band_row_bytes(std::min(TW, P.W), fx_max) : 64;
    int q_log2 = 0;
    while ((4 << q_log2) < TW) ++q_log2;
    const int groups = kConsumerThreads >> q_log2;
    const bool fast = !P.band_test && used == n_planes && P.out.vec4 && P.out.px_stride == 1;
    const int resident = fast ? 4 : 2;  // CTAs per SM the kernel is compiled for (__launch_bounds__)
    int TR = 32;
    while (TR > 1 && TR * 2 * rb_max > kStageBytesMax) TR /= 2;
    const int tiles_x = (P.W + TW - 1) / TW;
    // enough tiles to give every SM a couple of CTAs, but never fewer rows than the consumers process at once
    while (TR > groups && TR > 1 && static_cast<long long>(n_planes) * tiles_x * ((P.H + TR - 1) / TR) < 2LL * sm_count)
        TR /= 2;
    if (const char* e = std::getenv("CVGS_TMA_TR")) {  // tuning override (tests / profiling)
        const int v = std::atoi(e);
        if (v >= 1 && v <= 32 && v * 2 * rb_max <= kStageBytesMax) TR = v;
    }
    G.TW = TW;
    G.q_log2 = q_log2;
    G.groups = groups;
    G.TR = TR;
    G.tiles_x = tiles_x;
    G.tiles_y = (P.H + TR - 1) / TR;
    G.tiles_per_crop = G.tiles_x * G.tiles_y;
    const long long total = static_cast<long long>(n_planes) * G.tiles_per_crop;
    if (total > 0x7fffffffLL) return false;
    G.total_tiles = static_cast<int32_t>(total);
    G.stage_bytes = TR * 2 * rb_max;
    G.explicit_prescale = 0;
    const long long slots = static_cast<long long>(resident) * sm_count;
    if (total <= slots) {
        // small launch: one tile per CTA, every CTA resident at once, no ring
        G.grid = G.total_tiles;
        G.stages = 1;
    } else {
        // persistent: one CTA per slot, ring as deep as shared memory allows
        G.grid = static_cast<int32_t>(slots);
        const int per_cta = (227 * 1024) / resident - 1024 - 2 * kStagePad - 128;
        G.stages = std::max(1, std::min(kStages, per_cta / G.stage_bytes));
    }
    G.tiles_base = G.total_tiles / G.grid;
    G.tiles_rem = G.total_tiles % G.grid;
    return true;
}

// RN(1/d) in float.  1.0/d in double then rounded to float can be off by one ulp in rare double-rounding cases;
// d * r is exact in double (24 x 24 bits), so the candidate closest to 1 is picked exactly.
inline float correctly_rounded_reciprocal(float d) {
    const float r0 = static_cast<float>(1.0 / static_cast<double>(d));
    const float cand[3] = {std::nextafterf(r0, -INFINITY), r0, std::nextafterf(r0, INFINITY)};
    float best = r0;
    double best_err = INFINITY;
    for (float r : cand) {
        const double err = std::fabs(1.0 - static_cast<double>(d) * static_cast<double>(r));
        if (err < best_err) {
            best_err = err;
            best = r;
        }
    }
    return best;
}

// Chain for interpolated values: the 2^33 that undoes the tap/weight scaling is folded into the first op when
// that is exact for every input (MUL/FMA/DIV by a constant of moderate magnitude), else applied explicitly.
inline int scaled_program(const PreprocParams& P, TmaParams& K) {
    K.prog_img = P.prog;
    K.G.explicit_prescale = 1;
    if (P.prog.round_u8 || P.prog.n_ops == 0) return CH_GENERIC;
    auto moderate = [](float a) { return a == 0.f || (std::fabs(a) > 1e-20f && std::fabs(a) < 1e20f); };
    auto all_moderate = [&](const float* a, bool nonzero) {
        for (int c = 0; c < 3; ++c)
            if (!moderate(a[c]) || (nonzero && a[c] == 0.f)) return false;
        return true;
    };
    DevOp* ops = K.prog_img.ops;
    const int n = K.prog_img.n_ops;
    // canonical shape  v = fma(v, a, b) / d : [MUL|FMA|ADD] DIV  or  DIV alone.  MUL(a) == FMA(a, -0) and
    // ADD(b) == FMA(1, b) bit for bit (x + -0 == x for every x; x * 1 is exact).
    const bool first_lin = ops[0].kind == DOP_MUL || ops[0].kind == DOP_FMA || ops[0].kind == DOP_ADD;
    if ((n == 2 && first_lin && ops[1].kind == DOP_DIV) || (n == 1 && ops[0].kind == DOP_DIV)) {
        DevOp lin{};
        lin.kind = DOP_FMA;
        const DevOp div = ops[n - 1];
        for (int c = 0; c < 3; ++c) {
            lin.a[c] = n == 1 ? 1.0f : (ops[0].kind == DOP_ADD ? 1.0f : ops[0].a[c]);
            lin.b[c] = n == 1 ? -0.0f : (ops[0].kind == DOP_MUL ? -0.0f : (ops[0].kind == DOP_ADD ? ops[0].a[c] : ops[0].b[c]));
        }
        // fast exact division needs 2^-30 <= |d| <= 2^30 and |fma(x, a, b)| <= 2^60 for x in [0, 255]
        bool ok = all_moderate(lin.a, false);
        for (int c = 0; c < 3 && ok; ++c) {
            const double d = std::fabs(static_cast<double>(div.a[c]));
            const double vmax = 255.0 * std::fabs(static_cast<double>(lin.a[c])) + std::fabs(static_cast<double>(lin.b[c]));
            ok = std::isfinite(d) && d >= 9.313225746154785e-10 && d <= 1073741824.0 && vmax <= 1.0e18 &&
                 std::isfinite(lin.b[c]);
        }
        if (ok) {
            for (int c = 0; c < 3; ++c) {
                lin.a[c] *= kPreScale;
                K.div_rcp[c] = correctly_rounded_reciprocal(div.a[c]);
            }
            ops[0] = lin;
            ops[1] = div;
            K.prog_img.n_ops = 2;
            K.G.explicit_prescale = 0;
            return CH_FMA_DIV;
        }
        return CH_GENERIC;
    }
    DevOp& op = ops[0];
    if (op.kind == DOP_MUL || op.kind == DOP_FMA) {
        if (!all_moderate(op.a, false)) return CH_GENERIC;
        for (int c = 0; c < 3; ++c) op.a[c] *= kPreScale;
        K.G.explicit_prescale = 0;
    } else if (op.kind == DOP_DIV) {
        if (!all_moderate(op.a, true)) return CH_GENERIC;
        for (int c = 0; c < 3; ++c) op.a[c] /= kPreScale;
        K.G.explicit_prescale = 0;
    }
    return CH_GENERIC;
}

// Fill the TMA fields of a crop descriptor and encode its tensor map.
inline int tma_prepare_crop(DevCrop& c, const TmaGeom& G, int W, CUtensorMap* map) {
    const uintptr_t addr = reinterpret_cast<uintptr_t>(c.data);
    const int mis = static_cast<int>(addr & 15);
    const int rb = band_row_bytes(std::min(G.TW, W), c.fx);
    c.pad = rb | (mis << 16);
    const cuuint64_t dim[2] = {static_cast<cuuint64_t>((mis + 3LL * c.w + 7) / 8), static_cast<cuuint64_t>(c.h)};
    const cuuint64_t pitch = c.h > 1 ? static_cast<cuuint64_t>(c.pitch) : (dim[0] * 8 + 15) / 16 * 16;
    const cuuint64_t stride[1] = {pitch};
    const cuuint32_t box[2] = {static_cast<cuuint32_t>(rb / 8), 2};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = encode_tiled_fn()(map, CU_TENSOR_MAP_DATA_TYPE_UINT64, 2, reinterpret_cast<void*>(addr - mis), dim,
                                         stride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                         CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(CVGS_ERR_INVALID_VALUE, "cuTensorMapEncodeTiled failed (" + std::to_string((int)r) + ")");
    return CVGS_OK;
}

inline size_t tma_smem_bytes(const TmaGeom& G) {
    return static_cast<size_t>(G.stages) * G.stage_bytes + 2 * kStagePad + 128;
}

template <typename Table, int CHAIN, bool GEN>
inline int tma_launch_instance(const TmaParams& K, const Table& T, int device, cudaStream_t stream) {
    static thread_local size_t attr_set[64] = {};  // per device: dynamic shared memory opt-in already granted
    const size_t smem = tma_smem_bytes(K.G);
    const int slot = device & 63;
    auto kernel = preproc_tma_kernel<Table, CHAIN, GEN>;
    if (smem > attr_set[slot]) {
        const size_t want = std::max<size_t>(smem, 112 * 1024);
        CVGS_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(want)));
        attr_set[slot] = want;
    }
    kernel<<<K.G.grid, kTmaThreads, smem, stream>>>(K, T);
    count_launch();
    CVGS_CUDA(cudaGetLastError());
    return CVGS_OK;
}

template <typename Table>
inline int tma_launch_kernel(const TmaParams& K, const Table& T, int chain, int device, cudaStream_t stream) {
    const PreprocParams& P = K.P;
    const bool fast = !P.band_test && P.used == P.n_planes && P.out.vec4 && P.out.px_stride == 1;
    if (chain == CH_FMA_DIV)
        return fast ? tma_launch_instance<Table, CH_FMA_DIV, false>(K, T, device, stream)
                    : tma_launch_instance<Table, CH_FMA_DIV, true>(K, T, device, stream);
    return fast ? tma_launch_instance<Table, CH_GENERIC, false>(K, T, device, stream)
                : tma_launch_instance<Table, CH_GENERIC, true>(K, T, device, stream);
}

}  // namespace cvgs
