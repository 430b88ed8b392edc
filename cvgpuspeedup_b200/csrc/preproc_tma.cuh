// preproc_tma.cuh -- persistent, warp-specialised, TMA-staged kernel of the fused batch path (sm_100a).
//
// Same function as preproc_direct_kernel (preproc.cu) and as the reference instantiation of
// fk::launchTransformDPP_Kernel it replaces (reference fkl/include/fused_kernel/core/execution_model/
// data_parallel_patterns.cuh:157-197; BatchRead batch_operations.cuh:222-229; Resize resize.cuh:70-82,178-189;
// Interpolate interpolation.cuh:57-92; TensorSplit memory_operations.cuh:168-188), organised around the two
// budgets that bound the path on B200: issue slots and shared-memory wavefronts per output pixel.
//
//   * a CTA walks a contiguous range of tiles; a tile = TR output rows x TW = 32*NP output columns of one crop;
//   * the producer warp stages, per output row of the tile, the two source rows it taps with ONE
//     cp.async.bulk.tensor.2d (box = 2 rows x the tile's source span; byte-exact start coordinate through a
//     per-crop tensor map of 8-byte elements), completion on a per-stage mbarrier, and leaves the vertical taps of
//     the row (staged-row offsets + weights) next to the data, so consumers never convert or multiply for them;
//   * a consumer warp owns row PAIRS; lane l owns columns l, l+32, ... of the band (adjacent lanes read adjacent
//     source bytes: few shared-memory wavefronts; stores are full 128-byte lines).  The horizontal taps are
//     computed once per (crop, band), and the two rows of a pair ride in the two halves of packed-FP32
//     instructions (FMUL2 / FFMA2, sm_100): one issue slot per two roundings, each still IEEE round-to-nearest;
//   * u8 -> f32 is one PRMT: byte b placed at bits 16..23 of a float word is exactly b * 2^-133 (exponent
//     field 0 or 1, both scale 2^-149); the vertical weights carry 2^100 and the first op of the chain the
//     remaining 2^33 (kPreScale).  Scaling by powers of two commutes with every rounding below (no value
//     leaves the normal range), so results are bit-identical to the unscaled sequence;
//   * division by the launch constants is two operations (div_const.cpp), proven exact per divisor on the host.
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
#include "div_const.hpp"

namespace cvgs {

constexpr int kTmaParamCrops = 64;      // crops (descriptor + tensor map) that ride in the kernel parameters
constexpr int kConsumerWarps = 4;
constexpr int kConsumerThreads = kConsumerWarps * 32;
constexpr int kTmaThreads = kConsumerThreads + 32;  // + producer warp (the last one)
constexpr int kStages = 4;
constexpr int kMaxNP = 4;               // 32-column groups per band: a band is at most 128 output columns
constexpr int kStagePad = 128;          // bytes in front of / behind the stage ring (w[-1] / w[+1] over-reads)
constexpr int kMaxBoxBytes = 2048;      // 256 elements x 8 bytes
constexpr float kWeightScale = 1.2676506002282294e30f;  // 2^100
constexpr float kPreScale = 8589934592.0f;              // 2^33  = 2^133 / 2^100
constexpr uint32_t kRowFill = 0xFFFFFFF0u;  // RowInfo::offA: row lies outside the image band -> background
constexpr uint32_t kRowSkip = 0xFFFFFFFFu;  // RowInfo::offA: row is below the plane -> nothing to do

struct TmaGeom {
    int32_t NPB;             // 32-column groups per band (1..kMaxNP); band width TW = 32 * NPB
    int32_t TR;              // output rows per tile (even)
    int32_t tiles_x, tiles_y, tiles_per_crop, total_tiles;
    int32_t info_bytes;      // per stage: TR RowInfo records, rounded to 128 bytes
    int32_t stage_bytes;     // info_bytes + TR * 2 * max row bytes
    int32_t stages;          // depth of the stage ring actually used (1..kStages)
    int32_t resident;        // CTAs per SM the shared-memory footprint allows
    int32_t grid;            // CTAs; CTA b walks tiles [b*tiles_base + min(b, tiles_rem), ...)
    int32_t tiles_base, tiles_rem;
    int32_t explicit_prescale;  // 1: consumers multiply by 2^33 themselves (no op to fold it into)
    int32_t pdl_wait;        // 1: wait for the preceding kernel before the first global access (stream order);
                             // 0: the host proved independence, wait only before exiting (completion order)
};

struct TmaParams {
    PreprocParams P;         // P.prog = unscaled chain (background values), P.crops = device table or nullptr
    DevProgram prog_img;     // chain for interpolated values (2^33 folded into its first op)
    float zh[3], zl[3];      // CH_FMA_DIV: 1/d = zh + zl per source channel (div_const.cpp)
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

// Vertical taps of one output row of a tile, written by the producer next to the staged rows.
struct __align__(16) RowInfo {
    uint32_t offA;   // byte offset (from the stage's data block) of the staged upper source row, or kRowFill/kRowSkip
    uint32_t offB;   // ... of the lower source row (== offA when y2 is clamped, interpolation.cuh:73)
    float wy0, wy1;  // (y2 - sy) * 2^100, (sy - y1) * 2^100
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
__device__ __forceinline__ void tma_prefetch_2d(const CUtensorMap* map, int32_t c0, int32_t c1) {
    asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global.tile [%0, {%1, %2}];" ::"l"(reinterpret_cast<uint64_t>(map)), "r"(c0),
                 "r"(c1)
                 : "memory");
}
__device__ __forceinline__ uint32_t lds32(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ RowInfo lds_rowinfo(uint32_t addr) {
    RowInfo r;
    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r.offA), "=r"(r.offB), "=f"(r.wy0), "=f"(r.wy1)
                 : "r"(addr));
    return r;
}
__device__ __forceinline__ void sts_rowinfo(uint32_t addr, const RowInfo& r) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(r.offA), "r"(r.offB), "f"(r.wy0), "f"(r.wy1)
                 : "memory");
}
// Programmatic dependent launch (no-ops when the kernel was launched without the attribute).
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait_prior_grid() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

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
    const int tx0 = txi * (32 * G.NPB);
    b.xa = max(tx0, C.bx1);
    b.xe = min(min(tx0 + 32 * G.NPB, P.W) - 1, C.bx2);
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

// x / d for both halves with 1/d = zh + zl (div_const.cpp): FMUL2 + FFMA2.
__device__ __forceinline__ float2 div_by_const2(float2 x, float zh, float zl) {
    const float2 u = __fmul2_rn(x, make_float2(zl, zl));
    return __ffma2_rn(x, make_float2(zh, zh), u);
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
    // n tiles further down the same band (n <= tiles_y - tyi)
    __device__ __forceinline__ void skip(const TmaGeom& G, int n) {
        left -= n;
        tyi += n;
        if (tyi == G.tiles_y) {
            tyi = 0;
            if (++txi == G.tiles_x) {
                txi = 0;
                ++z;
            }
        }
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

// One output column of a row pair: taps of both rows from their staged source rows -> 3 interpolated channels,
// row r in .x and row r+1 in .y (scaled by 2^-33).  Arithmetic per half = Interpolate<INTER_LINEAR>::exec in the
// order nvcc emits for the reference: FMUL(p10*w10), FFMA(p00,w00), FFMA(p01,w01), FFMA(p11,w11).
__device__ __forceinline__ void gather_pair(uint32_t A0, uint32_t B0, uint32_t A1, uint32_t B1, int shl, int shr, bool edge,
                                            float wx0, float wx1, float2 wy0, float2 wy1, float2 (&v)[3]) {
    const uint32_t am0 = lds32(A0 - 4), a00 = lds32(A0), a01 = lds32(A0 + 4);
    const uint32_t bm0 = lds32(B0 - 4), b00 = lds32(B0), b01 = lds32(B0 + 4);
    const uint32_t am1 = lds32(A1 - 4), a10 = lds32(A1), a11 = lds32(A1 + 4);
    const uint32_t bm1 = lds32(B1 - 4), b10 = lds32(B1), b11 = lds32(B1 + 4);
    // left pixel in bytes 0..2 (clamped shift: 32 = the word itself), right pixel in bytes 0..2
    const uint32_t al0 = __funnelshift_rc(am0, a00, shl), bl0 = __funnelshift_rc(bm0, b00, shl);
    const uint32_t al1 = __funnelshift_rc(am1, a10, shl), bl1 = __funnelshift_rc(bm1, b10, shl);
    uint32_t ar0 = __funnelshift_r(a00, a01, shr), br0 = __funnelshift_r(b00, b01, shr);
    uint32_t ar1 = __funnelshift_r(a10, a11, shr), br1 = __funnelshift_r(b10, b11, shr);
    if (edge) {  // x2_read == x1 (interpolation.cuh:72): the right tap is the left pixel again
        ar0 = al0;
        br0 = bl0;
        ar1 = al1;
        br1 = bl1;
    }
    const float2 w00 = __fmul2_rn(make_float2(wx0, wx0), wy0), w10 = __fmul2_rn(make_float2(wx1, wx1), wy0);
    const float2 w01 = __fmul2_rn(make_float2(wx0, wx0), wy1), w11 = __fmul2_rn(make_float2(wx1, wx1), wy1);
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        float2 t = __fmul2_rn(make_float2(u8_scaled(ar0, c), u8_scaled(ar1, c)), w10);
        t = __ffma2_rn(make_float2(u8_scaled(al0, c), u8_scaled(al1, c)), w00, t);
        t = __ffma2_rn(make_float2(u8_scaled(bl0, c), u8_scaled(bl1, c)), w01, t);
        v[c] = __ffma2_rn(make_float2(u8_scaled(br0, c), u8_scaled(br1, c)), w11, t);
    }
}

// The normalised chain on a row pair (same semantics as apply_program, two values per instruction where the
// hardware has a packed form).
__device__ __forceinline__ void apply_program_pair(const DevProgram& prog, float2 (&v)[3]) {
    if (prog.round_u8) {
#pragma unroll
        for (int c = 0; c < 3; ++c) v[c] = make_float2(round_sat_u8(v[c].x), round_sat_u8(v[c].y));
    }
    for (int i = 0; i < prog.n_ops; ++i) {
        const DevOp& op = prog.ops[i];
        switch (op.kind) {
            case DOP_FMA:
#pragma unroll
                for (int c = 0; c < 3; ++c) v[c] = __ffma2_rn(v[c], make_float2(op.a[c], op.a[c]), make_float2(op.b[c], op.b[c]));
                break;
            case DOP_MUL:
#pragma unroll
                for (int c = 0; c < 3; ++c) v[c] = __fmul2_rn(v[c], make_float2(op.a[c], op.a[c]));
                break;
            case DOP_ADD:
#pragma unroll
                for (int c = 0; c < 3; ++c) v[c] = make_float2(__fadd_rn(v[c].x, op.a[c]), __fadd_rn(v[c].y, op.a[c]));
                break;
            case DOP_DIV:
#pragma unroll
                for (int c = 0; c < 3; ++c) v[c] = make_float2(__fdiv_rn(v[c].x, op.a[c]), __fdiv_rn(v[c].y, op.a[c]));
                break;
            default:
                break;
        }
    }
}

// GEN = false: the common geometry -- IGNORE_AR, every plane used, planar output.
// GEN = true : aspect-ratio bands, unused planes, packed outputs.
template <typename Table, int CHAIN, bool GEN>
__global__ void __launch_bounds__(kTmaThreads, 4)
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

    pdl_launch_dependents();  // the next kernel of the stream may start its prologue now

    // stage ring, 128-byte aligned, with kStagePad bytes of slack on both sides
    uint32_t ring = ((smem_u32(smem_raw) + 127u) & ~127u) + kStagePad;
    asm volatile("" : "+r"(ring));  // keep it in a register (the compiler would re-derive it per use)

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
        if (G.pdl_wait) pdl_wait_prior_grid();  // source images may be written by the preceding kernel
        int stage = 0;
        uint32_t phase = 0;
        if (K.maps) {
            // Tensor maps that reached global memory through a host copy must be acquired for the TMA proxy once
            // per CTA and map before their first use (not per load: the fence also drops the descriptor cache).
            const int t_last = blockIdx.x * G.tiles_base + min((int)blockIdx.x, G.tiles_rem) + tc.left - 1;
            const int z_last = t_last / G.tiles_per_crop;
            for (int z = tc.z; z <= z_last; ++z)
                asm volatile("fence.proxy.tensormap::generic.acquire.gpu [%0], 128;" ::"l"(reinterpret_cast<uint64_t>(K.maps + z))
                             : "memory");
        }
        // L2 prefetch cursor: the boxes of the tile `nstages` ahead are requested into L2 while their shared-memory
        // slot is still occupied, so that the real load later pays L2 latency instead of DRAM latency
        TileCursor pf = tc;
        for (int i = 0; i < nstages && pf.left > 0; ++i) pf.next(G);
        for (; tc.left > 0; tc.next(G)) {
            if (pf.left > 0) {
                const int y = pf.tyi * G.TR + lane;
                if (lane < G.TR && y < P.H && (!GEN || pf.z < P.used)) {
                    const DevCrop& C = tma_crop_of<Table>(K, T, pf.z);
                    const BandOrigin b = band_origin(P, G, C, pf.txi);
                    if (y >= C.by1 && y <= C.by2 && b.xa <= b.xe)
                        tma_prefetch_2d(tma_map_of<Table>(K, T, pf.z), b.c0, axis_tap(y - C.by1, C.fy).i1);
                }
                pf.next(G);
            }
            // everything that does not touch the stage is computed before waiting for it: once the consumers
            // release the slot only the RowInfo store, the arrive and the TMA issue remain
            const uint32_t full = smem_u32(&bar_full[stage]);
            const uint32_t sinfo = ring + stage * G.stage_bytes;
            bool issue = false;
            int i1 = 0, c0 = 0, rb = 0;
            RowInfo ri;
            ri.offA = kRowSkip;
            ri.offB = 0;
            ri.wy0 = ri.wy1 = 0.f;
            if (lane < G.TR) {
                const int y = tc.tyi * G.TR + lane;
                if (y < P.H) {
                    ri.offA = kRowFill;
                    if (!GEN || tc.z < P.used) {
                        const DevCrop& C = tma_crop_of<Table>(K, T, tc.z);
                        const BandOrigin b = band_origin(P, G, C, tc.txi);
                        rb = crop_row_bytes(C);
                        if (y >= C.by1 && y <= C.by2 && b.xa <= b.xe) {
                            issue = true;
                            const AxisTap t = axis_tap(y - C.by1, C.fy);
                            i1 = t.i1;
                            c0 = b.c0;
                            ri.offA = (uint32_t)(lane * 2 * rb);
                            ri.offB = ri.offA + ((t.i1 + 1 > C.h - 1) ? 0u : (uint32_t)rb);
                            ri.wy0 = __fmul_rn(t.w0, kWeightScale);
                            ri.wy1 = __fmul_rn(t.w1, kWeightScale);
                        }
                    }
                }
            }
            const unsigned m = __ballot_sync(0xffffffffu, issue);
            const int rb_all = __shfl_sync(0xffffffffu, rb, m ? (__ffs(m) - 1) : 0);
            const uint32_t tx_bytes = (uint32_t)(__popc(m) * 2 * rb_all);
            const CUtensorMap* map = tma_map_of<Table>(K, T, tc.z);
            const uint32_t dst = sinfo + G.info_bytes + lane * 2 * rb;

            mbar_wait(smem_u32(&bar_empty[stage]), phase ^ 1u);
            if (lane < G.TR) sts_rowinfo(sinfo + lane * (uint32_t)sizeof(RowInfo), ri);
            __syncwarp();  // the RowInfo stores of all lanes are ordered before lane 0's (releasing) arrive
            if (lane == 0) mbar_arrive_expect_tx(full, tx_bytes);
            if (issue) tma_load_2d(dst, map, c0, i1, full);
            if (++stage == nstages) {
                stage = 0;
                phase ^= 1u;
            }
        }
        if (!G.pdl_wait) pdl_wait_prior_grid();  // never complete before the preceding kernel has
        return;
    }

    // ========================================= consumers =========================================
    const int W = P.W;
    const int TW = 32 * G.NPB;

    // chain constants of the specialised shape v = fma(v, ca, cb) / cd  (source-channel order)
    float ca[3], cb[3], zh[3], zl[3];
    if (CHAIN == CH_FMA_DIV) {
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            ca[c] = K.prog_img.ops[0].a[c];
            cb[c] = K.prog_img.ops[0].b[c];
            zh[c] = K.zh[c];
            zl[c] = K.zl[c];
        }
    }
    // chain(background): value of planes z >= used and of pixels outside the aspect-ratio band
    float vb[1][3] = {{P.bg[0], P.bg[1], P.bg[2]}};
    if (GEN) apply_program<1>(P.prog, vb);

    // plane offsets (floats) of the three source channels
    const long long oc0 = (long long)P.prog.dst_chan[0] * P.out.c_stride;
    const long long oc1 = (long long)P.prog.dst_chan[1] * P.out.c_stride;
    const long long oc2 = (long long)P.prog.dst_chan[2] * P.out.c_stride;
    const int pxs = GEN ? P.out.px_stride : 1;
    const int row_step = W * pxs;            // floats between vertically adjacent pixels
    const int tile_step = G.TR * row_step;   // ... between the first rows of vertically adjacent tiles
    const uint32_t info_bytes = (uint32_t)G.info_bytes, stage_bytes = (uint32_t)G.stage_bytes;
    const int half_rows = G.TR >> 1;

    int stage = 0;
    uint32_t phase = 0;
    // From here on this thread stores to global memory: order it after the preceding kernel unless the host
    // proved the two independent (then only completion is ordered, at the end).
    if (G.pdl_wait) pdl_wait_prior_grid();

    while (tc.left > 0) {
        // ---------------- horizontal state of this lane for the band (z, txi): column p is tx0 + lane + 32 p ------
        const int z = tc.z;
        const int tx0 = tc.txi * TW;
        const int np = (min(TW, W - tx0) + 31) >> 5;
        int32_t off[kMaxNP], shl[kMaxNP], shr[kMaxNP];
        float wxa[kMaxNP], wxb[kMaxNP];
        bool edge[kMaxNP], img[kMaxNP], inw[kMaxNP];  // right tap clamped / column receives image data / column < W
        {
            const bool active = !GEN || z < P.used;
            BandOrigin b;
            b.xa = 1;
            b.xe = 0;
            b.c0 = b.origin = 0;
            float fx = 1.f;
            int bx1 = 0, wm1 = 0;
            if (active) {
                const DevCrop& C = tma_crop_of<Table>(K, T, z);
                b = band_origin(P, G, C, tc.txi);
                fx = C.fx;
                bx1 = C.bx1;
                wm1 = C.w - 1;
            }
#pragma unroll
            for (int p = 0; p < kMaxNP; ++p) {
                const int x = tx0 + lane + 32 * p;
                inw[p] = x < W;
                img[p] = inw[p] && x >= b.xa && x <= b.xe;
                const AxisTap t = axis_tap((img[p] ? x : b.xa) - bx1, fx);
                wxa[p] = t.w0;
                wxb[p] = t.w1;
                edge[p] = t.i1 + 1 > wm1;
                const int o = 3 * t.i1 - b.origin;
                off[p] = ((o + 3) >> 2) * 4;
                shl[p] = (o & 3) ? (o & 3) * 8 : 32;
                shr[p] = ((o + 3) & 3) * 8;
            }
        }
        // first pixel of this lane in the three channel planes of plane z
        float* bp0 = P.out.base + ((long long)z * P.out.z_stride + (long long)(tx0 + lane) * pxs);
        float* bp1 = bp0 + oc1;
        float* bp2 = bp0 + oc2;
        bp0 += oc0;
        asm volatile("" : "+l"(bp0), "+l"(bp1), "+l"(bp2));

        const int ntiles = min(tc.left, G.tiles_y - tc.tyi);  // tiles of this band inside the CTA's range
        int tile_off = tc.tyi * tile_step;                    // in-plane offset (floats) of the tile's first row
        for (int t = 0; t < ntiles; ++t, tile_off += tile_step) {
            mbar_wait(smem_u32(&bar_full[stage]), phase);
            const uint32_t sinfo = ring + stage * stage_bytes;
            const uint32_t sdata = sinfo + info_bytes;

            for (int j = warp; j < half_rows; j += kConsumerWarps) {
                const RowInfo r0 = lds_rowinfo(sinfo + 2 * j * (uint32_t)sizeof(RowInfo));
                const RowInfo r1 = lds_rowinfo(sinfo + (2 * j + 1) * (uint32_t)sizeof(RowInfo));
                if (r0.offA == kRowSkip) break;  // rows are ascending: nothing below either
                const bool st1 = r1.offA != kRowSkip;
                const bool im0 = !GEN || r0.offA < kRowFill, im1 = r1.offA < kRowFill;
                // a row without image data borrows the other row's taps (its values are replaced / not stored)
                uint32_t aA0, aB0, aA1, aB1;
                float2 wy0, wy1;
                if (GEN) {
                    aA0 = sdata + (im0 ? r0.offA : r1.offA), aB0 = sdata + (im0 ? r0.offB : r1.offB);
                    wy0.x = im0 ? r0.wy0 : r1.wy0, wy1.x = im0 ? r0.wy1 : r1.wy1;
                } else {
                    aA0 = sdata + r0.offA, aB0 = sdata + r0.offB;
                    wy0.x = r0.wy0, wy1.x = r0.wy1;
                }
                aA1 = sdata + (im1 ? r1.offA : r0.offA), aB1 = sdata + (im1 ? r1.offB : r0.offB);
                wy0.y = im1 ? r1.wy0 : r0.wy0, wy1.y = im1 ? r1.wy1 : r0.wy1;
                // row pointers of the pair for the three channels; opaque to the compiler so that they are kept in
                // registers instead of being re-derived in front of every store
                const int ro = tile_off + 2 * j * row_step;
                float* s0 = bp0 + ro;
                float* s1 = bp1 + ro;
                float* s2 = bp2 + ro;
                float* t0 = s0 + row_step;
                float* t1 = s1 + row_step;
                float* t2 = s2 + row_step;
                asm volatile("" : "+l"(s0), "+l"(s1), "+l"(s2), "+l"(t0), "+l"(t1), "+l"(t2));
                asm volatile("" : "+r"(aA0), "+r"(aB0), "+r"(aA1), "+r"(aB1), "+f"(wy0.x), "+f"(wy0.y), "+f"(wy1.x), "+f"(wy1.y));
#pragma unroll
                for (int p = 0; p < kMaxNP; ++p) {
                    if (p < np) {
                        float2 v[3];
                        if (!GEN || im0 || im1) {
                            gather_pair(aA0 + off[p], aB0 + off[p], aA1 + off[p], aB1 + off[p], shl[p], shr[p], edge[p],
                                        wxa[p], wxb[p], wy0, wy1, v);
                            if (CHAIN == CH_FMA_DIV) {
#pragma unroll
                                for (int c = 0; c < 3; ++c) {
                                    v[c] = __ffma2_rn(v[c], make_float2(ca[c], ca[c]), make_float2(cb[c], cb[c]));
                                    v[c] = div_by_const2(v[c], zh[c], zl[c]);
                                }
                            } else {
                                if (G.explicit_prescale) {
#pragma unroll
                                    for (int c = 0; c < 3; ++c) v[c] = __fmul2_rn(v[c], make_float2(kPreScale, kPreScale));
                                }
                                apply_program_pair(K.prog_img, v);
                            }
                        }
                        if (GEN) {
#pragma unroll
                            for (int c = 0; c < 3; ++c) {
                                if (!(im0 && img[p])) v[c].x = vb[0][c];
                                if (!(im1 && img[p])) v[c].y = vb[0][c];
                            }
                        }
                        if (inw[p]) {
                            const int q = 32 * p * pxs;
                            st_cs_f32(s0 + q, v[0].x);
                            st_cs_f32(s1 + q, v[1].x);
                            st_cs_f32(s2 + q, v[2].x);
                            if (st1) {
                                st_cs_f32(t0 + q, v[0].y);
                                st_cs_f32(t1 + q, v[1].y);
                                st_cs_f32(t2 + q, v[2].y);
                            }
                        }
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
        tc.skip(G, ntiles);
    }
    if (!G.pdl_wait) pdl_wait_prior_grid();  // never complete before the preceding kernel has
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
    if (static_cast<long long>(P.W) * P.H * 3 * std::max<long long>(1, std::abs(P.out.px_stride)) > 0x3fffffffLL)
        return false;  // in-plane offsets are 32-bit in the kernel
    int NPB = std::min(kMaxNP, (P.W + 31) / 32);
    auto need = [&](int npb) { return used > 0 ? band_row_bytes(std::min(32 * npb, P.W), fx_max) : 64; };
    while (NPB > 1 && need(NPB) > kMaxBoxBytes) --NPB;
    if (need(NPB) > kMaxBoxBytes) return false;  // extreme down-scale: direct kernel
    const int rb_max = need(NPB);
    const int TW = 32 * NPB;
    const int tiles_x = (P.W + TW - 1) / TW;
    const int smem_sm = 227 * 1024;
    auto stage_of = [&](int tr) { return (tr * static_cast<int>(sizeof(RowInfo)) + 127) / 128 * 128 + tr * 2 * rb_max; };
    auto resident_of = [&](int tr, int stages) {
        const int per_cta = stages * stage_of(tr) + 2 * kStagePad + 128 + 1024 /*static + reserved*/;
        return std::min(4, smem_sm / per_cta);
    };
    // rows per tile: first keep as many CTAs per SM as the smallest tile allows (occupancy hides the latency of the
    // dependent FP chain), then take the tallest tile that still gives every CTA slot at least one tile
    int TR = 8;
    const int best_resident = resident_of(8, 2);
    for (int tr : {32, 16}) {
        if (resident_of(tr, 2) < best_resident) continue;
        const long long tiles = static_cast<long long>(n_planes) * tiles_x * ((P.H + tr - 1) / tr);
        if (tiles >= static_cast<long long>(best_resident) * sm_count) {
            TR = tr;
            break;
        }
    }
    if (const char* e = std::getenv("CVGS_TMA_TR")) {  // tuning override (tests / profiling)
        const int v = std::atoi(e);
        if ((v == 8 || v == 16 || v == 32) && resident_of(v, 1) >= 1) TR = v;
    }
    if (resident_of(TR, 1) < 1) return false;
    G.NPB = NPB;
    G.TR = TR;
    G.tiles_x = tiles_x;
    G.tiles_y = (P.H + TR - 1) / TR;
    G.tiles_per_crop = G.tiles_x * G.tiles_y;
    const long long total = static_cast<long long>(n_planes) * G.tiles_per_crop;
    if (total > 0x7fffffffLL) return false;
    G.total_tiles = static_cast<int32_t>(total);
    G.info_bytes = (TR * static_cast<int>(sizeof(RowInfo)) + 127) / 128 * 128;
    G.stage_bytes = stage_of(TR);
    G.explicit_prescale = 0;
    G.pdl_wait = 1;
    int stages = 2;
    int resident = resident_of(TR, stages);
    if (resident < 1) {
        stages = 1;
        resident = resident_of(TR, 1);
    }
    // deeper ring when it costs no residency
    while (stages < kStages && resident_of(TR, stages + 1) >= resident) ++stages;
    const long long slots = static_cast<long long>(resident) * sm_count;
    G.resident = resident;
    if (total <= slots) {
        // small launch: one tile per CTA, every CTA resident at once, no ring
        G.grid = G.total_tiles;
        G.stages = 1;
    } else {
        // persistent: one CTA per slot
        G.grid = static_cast<int32_t>(slots);
        G.stages = stages;
    }
    G.tiles_base = G.total_tiles / G.grid;
    G.tiles_rem = G.total_tiles % G.grid;
    return true;
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
        // The two-operation division is proven for normal numerators (div_const.cpp).  With 2^-24 <= |a|, |b|, |d|
        // <= 2^24 (b may be zero) a non-zero fma(x, a, b) of an interpolated x in {0} U [2^-46, 255] has magnitude
        // in [2^-73, 2^33]: every intermediate stays normal.  Zeros are handled through the signs of zh / zl.
        bool ok = true;
        DivConst dc[3];
        for (int c = 0; c < 3 && ok; ++c) {
            const float aa = std::fabs(lin.a[c]), ab = std::fabs(lin.b[c]);
            ok = std::isfinite(aa) && aa >= 5.9604644775390625e-08f && aa <= 16777216.0f &&
                 (ab == 0.f || (std::isfinite(ab) && ab >= 5.9604644775390625e-08f && ab <= 16777216.0f));
            if (!ok) break;
            dc[c] = div_const_prepare(div.a[c]);
            // fma(x, a, b) is -0 only for x = +0 with a < 0 and b = -0; +0 whenever the sum is an exact zero otherwise
            const bool neg_zero_possible = ab == 0.f && std::signbit(lin.b[c]) && std::signbit(lin.a[c]);
            ok = dc[c].exact && div_const_pos_zero_ok(dc[c]) && (!neg_zero_possible || div_const_neg_zero_ok(dc[c]));
        }
        if (ok) {
            for (int c = 0; c < 3; ++c) {
                lin.a[c] *= kPreScale;
                K.zh[c] = dc[c].zh;
                K.zl[c] = dc[c].zl;
            }
            ops[0] = lin;
            ops[1] = div;
            K.prog_img.n_ops = 2;
            K.G.explicit_prescale = 0;
            return CH_FMA_DIV;
        }
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
    const int rb = band_row_bytes(std::min(32 * G.NPB, W), c.fx);
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
    // Programmatic dependent launch: the kernel may be scheduled while the preceding kernel of the stream drains;
    // it orders itself with griddepcontrol.wait (G.pdl_wait), so stream semantics are unchanged.
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(static_cast<unsigned>(K.G.grid));
    cfg.blockDim = dim3(kTmaThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    CVGS_CUDA(cudaLaunchKernelEx(&cfg, kernel, K, T));
    count_launch();
    return CVGS_OK;
}

template <typename Table>
inline int tma_launch_kernel(const TmaParams& K, const Table& T, int chain, int device, cudaStream_t stream) {
    const PreprocParams& P = K.P;
    const bool fast = !P.band_test && P.used == P.n_planes && P.out.px_stride == 1;
    if (chain == CH_FMA_DIV)
        return fast ? tma_launch_instance<Table, CH_FMA_DIV, false>(K, T, device, stream)
                    : tma_launch_instance<Table, CH_FMA_DIV, true>(K, T, device, stream);
    return fast ? tma_launch_instance<Table, CH_GENERIC, false>(K, T, device, stream)
                : tma_launch_instance<Table, CH_GENERIC, true>(K, T, device, stream);
}

}  // namespace cvgs
