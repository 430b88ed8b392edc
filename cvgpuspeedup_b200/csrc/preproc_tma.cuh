// preproc_tma.cuh -- persistent, TMA-staged kernel of the fused batch path (sm_100a).
//
// Same function as preproc_direct_kernel (preproc.cu) and as the reference instantiation of
// fk::launchTransformDPP_Kernel it replaces (reference fkl/include/fused_kernel/core/execution_model/
// data_parallel_patterns.cuh:157-197; BatchRead batch_operations.cuh:222-229; Resize resize.cuh:70-82,178-189;
// Interpolate interpolation.cuh:57-92; TensorSplit memory_operations.cuh:168-188), organised around the budgets
// that bound the path on B200: issue slots and shared-memory wavefronts per output pixel, and the
// latency of the staging pipeline.
//
//   * the unit of work is an ITEM = one pair of output rows (2j, 2j+1) x one band of up to 128 output columns of one
//     crop; every warp owns a contiguous range of items and feeds itself: it stages the four source rows its next
//     items tap with cp.async.bulk.tensor.2d (one box of 2 source rows per output row, byte-exact start coordinate
//     through a per-crop tensor map of 8-byte elements) into its private ring of shared-memory slots, completion
//     on a per-slot mbarrier.  No producer warp, no cross-warp synchronisation: a slot is refilled by the warp
//     that just finished reading it;
//   * lane l owns columns l, l+32, ... of the band (adjacent lanes read adjacent source bytes: few shared-memory
//     wavefronts; stores are full 128-byte lines).  The horizontal taps are computed once per (crop, band), the
//     vertical ones once per item by two lanes; the two rows of the pair ride in the two halves of packed-FP32
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
#include <type_traits>

#include "../../include/cvgs_b200.h"
#include "cvgs_device.cuh"
#include "cvgs_runtime.hpp"
#include "div_const.hpp"

namespace cvgs {

constexpr int kTmaParamCrops = 64;      // crops (descriptor + tensor map) that ride in the kernel parameters
constexpr int kWarps = 4;               // warps per CTA, each with its own slot ring
constexpr int kTmaThreads = kWarps * 32;
constexpr int kMaxSlots = 4;            // slots per warp
#ifndef CVGS_MAX_RESIDENT  // diagnostic builds (scripts/diag_build.sh): 4 lets the compiler use 128 registers
#define CVGS_MAX_RESIDENT 5
#endif
constexpr int kMaxResident = CVGS_MAX_RESIDENT;  // CTAs per SM the kernel is compiled for (5: <= 102 registers per thread)
constexpr int kMaxNP = 4;               // 32-column groups per band: a band is at most 128 output columns
constexpr int kSlotHeader = 128;        // per slot: room for the w[-1] over-read in front of the staged rows
constexpr int kRingPad = 128;           // bytes behind the last slot (w[+1] over-read)
constexpr int kMaxBoxBytes = 2048;      // 256 elements x 8 bytes
constexpr float kWeightScale = 1.2676506002282294e30f;  // 2^100
constexpr float kPreScale = 8589934592.0f;              // 2^33  = 2^133 / 2^100   (8-bit sources)
constexpr float kPreScale16 = 2199023255552.0f;         // 2^41  = 2^141 / 2^100   (16-bit sources: halfword at mantissa bits 8..23)
constexpr uint32_t kRowFill = 0xFFFFFFF0u;  // RowTap::a: row lies outside the image band -> background
constexpr uint32_t kRowSkip = 0xFFFFFFFFu;  // RowTap::a: row is below the plane (or past the warp's range) -> nothing to do

// Division of n < 2^31 by a launch constant: q = umulhi(n, mul) >> shr (mul == 0: divisor 1).  Exact on that range
// (round-up method: p = 31 + ceil(log2 d), mul = ceil(2^p / d)); the warps divide item and cost indices by the plan's
// constants in their prologues, where a 32-bit hardware-less division costs ~40 instructions each.
struct FastDiv {
    uint32_t mul, shr;
};
inline FastDiv fast_div_make(uint32_t d) {
    FastDiv f{0u, 0u};
    if (d > 1) {
        uint32_t l = 0;
        while ((1ull << l) < d) ++l;
        const uint32_t p = 31 + l;
        f.mul = static_cast<uint32_t>(((1ull << p) + d - 1) / d);
        f.shr = p - 32;
    }
    return f;
}
__host__ __device__ __forceinline__ uint32_t fast_div(uint32_t n, FastDiv f) {
#ifdef __CUDA_ARCH__
    return f.mul ? __umulhi(n, f.mul) >> f.shr : n;
#else
    return f.mul ? static_cast<uint32_t>((static_cast<unsigned long long>(n) * f.mul) >> 32) >> f.shr : n;
#endif
}

struct TmaGeom {
    int32_t NPB;             // 32-column groups per band (1..kMaxNP); band width TW = 32 * NPB
    int32_t HP;              // row pairs per plane = ceil(H / 2)
    int32_t tiles_x;         // bands per plane
    int32_t items_per_crop;  // tiles_x * HP
    int32_t total_items;
    int32_t slot_bytes;      // kSlotHeader + 4 * max row bytes
    int32_t slots;           // ring depth per warp (1..kMaxSlots)
    int32_t resident;        // CTAs per SM the shared-memory footprint allows
    int32_t grid;            // CTAs; warp g = blockIdx.x * kWarps + warp walks a contiguous range of items whose
                             // total cost (32-column groups) is 1/(grid*kWarps) of the launch (ItemCursor::init)
    int32_t np_last;         // 32-column groups of the last band of a plane (the others have NPB)
    int32_t w_full, w_crop;  // cost (column groups) of the full-width bands of a plane / of a whole plane
    int32_t share_q, share_r;  // total cost = share_q * (grid * kWarps) + share_r
    int32_t explicit_prescale;  // 1: the kernel multiplies by `prescale` itself (no op to fold it into)
    float prescale;          // what undoes the tap / weight scaling: 2^33 for 8-bit sources, 2^41 for 16-bit ones
    int32_t pdl_wait;        // 1: wait for the preceding kernel before the first global access (stream order);
                             // 0: the host proved independence, wait only before exiting (completion order)
    FastDiv d_w_crop, d_NPB, d_np_last, d_items_per_crop, d_HP;  // divisions by the fields of those names
};

constexpr int kMaxDest = 8;  // replicas of the output tensor one launch can write (this GPU's + peer-mapped ones)
struct TmaParams {
    PreprocParams P;         // P.prog = unscaled chain (background values), P.crops = device table or nullptr
    DevProgram prog_img;     // chain for interpolated values (2^33 folded into its first op)
    float zh[4], zl[4];      // CH_FMA_DIV: 1/d = zh + zl per source channel (div_const.cpp)
    // CH_GRAY (gray_setup): PRMT selectors that fetch the conversion's first / second / third operand channel from a pixel
    // word (the gather then delivers v[0..2] in operand order: no selects in the loop), their coefficients x 2^33, whether
    // the products are rounded on their own (CVGS_FP_SEPARATE), and the ops behind the conversion in the canonical form
    // g = fma(g, ga, gb) / d (two-operation division; gmode 2) or left to the interpreter (gmode 3)
    uint32_t gsel[3];
    float gk[3];
    int32_t gsep, gmode;
    float ga, gb, gzh, gzl;
    float alpha;             // CH_*_ALPHA: value of the alpha plane
    long long alpha_delta;   //             floats from the plane of source channel 0 to the alpha plane
    TmaGeom G;
    const CUtensorMap* maps; // device table (nullptr when the maps ride in the kernel parameters)
    // PEER instantiation (cvgs_b200_preproc_launch_replicated): every value is stored n_dest times, at
    // P.out.base + dest_delta[d] (floats; dest_delta[0] = 0) -- the same tensor on this GPU and, through peer-mapped
    // pointers, on the other GPUs of the box, so that the gather of BASELINE config 5 rides on the kernel's own stores
    int32_t n_dest;
    long long dest_delta[kMaxDest];
};

// Descriptors that ride in the kernel parameters: NMAPS tensor maps + up to kTmaParamCrops crops.
// One map per crop (NMAPS = kTmaParamCrops) when nothing is known about the memory around a crop; a handful of
// per-image maps (NMAPS = kTmaImageMaps) when the caller names the parent images (cvgs_b200_preproc_launch_ex).
constexpr int kTmaImageMaps = 16;
constexpr int kTmaImageCrops = 256;     // image mode: batches up to this size still ride in the parameters (13 KB)
template <int NMAPS, int NCROPS>
struct alignas(64) TmaParamTableT {
    CUtensorMap m[NMAPS];
    DevCrop c[NCROPS];
};
using TmaParamTable = TmaParamTableT<kTmaParamCrops, kTmaParamCrops>;
using TmaImageTable = TmaParamTableT<kTmaImageMaps, kTmaParamCrops>;
using TmaImageTableL = TmaParamTableT<kTmaImageMaps, kTmaImageCrops>;
struct TmaNoTable {
    int32_t unused;
};
// Coalesced frame groups (cvgs_b200_preproc_launch_sequence_ex): the crops of up to kMultiGroups independent argument
// sets -- each with its own parent frame and its own output tensor -- ride in ONE launch.  Descriptors in the kernel
// parameters, tensor maps in the device-resident cache (TmaParams::maps, DevMapCache below).
constexpr int kMultiCrops = 928;
constexpr int kMultiGroups = 32;
// Shared launches take the common geometry only (IGNORE_AR: the image band is the whole plane), so a crop needs 32 bytes
// instead of DevCrop's 48 -- 928 crops (18 frames of 50) instead of 600 fit the 32 KB of kernel parameters.
struct __align__(16) DevCropC {
    int32_t xb, y0;   // where the ROI sits inside its tensor map (DevCrop::m)
    int32_t w, h;     // source size in pixels
    float fx, fy;     // src_conv_factors
    int32_t pad;      // bits 0..15 staged row bytes, bits 16..31 index of the tensor map
    int32_t group;    // index of the crop's argument set
};
static_assert(sizeof(DevCropC) == 32, "DevCropC layout");
struct alignas(64) TmaMultiTable {
    DevCropC c[kMultiCrops];
    float* out_base[kMultiGroups];   // output tensor of group g
    int32_t z_first[kMultiGroups];   // launch-wide plane index of the group's first crop
};

// Vertical taps of one output row, computed 16 rows (8 items) at a time by 16 lanes of the warp that owns the
// items and kept in a per-warp table; both the staging of an item and its computation read them from there.
struct __align__(16) RowTap {
    uint32_t a;      // kRowSkip / kRowFill, or y1 (first source row, relative to the crop) | kTapClamp when y2 == y1
    float wy0, wy1;  // (y2 - sy) * 2^100, (sy - y1) * 2^100
    uint32_t pad;
};
constexpr uint32_t kTapClamp = 1u << 30;   // interpolation.cuh:73: y2_read == y1, both taps read the same source row
constexpr int kTapBlock = 8;               // items per table block; the table holds two blocks (consumers lag <= kMaxSlots)
// What the warp's k-th item needs, derived once per kTapBlock items by the lanes in parallel (compute_taps) and read
// back by the consumer loop (first 32 bytes) and by the staging code (last 16 bytes): nothing about an item's rows
// is recomputed in the per-item code.
struct __align__(16) ItemRec {
    uint32_t a0;     // shared-memory offsets from the slot's data of the two source rows output row 0 taps: lo16 upper, hi16
                     // lower (multiples of 64: staged row bytes are).  Bits 0..3 of the word are flags (kRec*)
    uint32_t a1;     // same for output row 1 (a row without image data borrows the other row's: its values are replaced / not stored)
    int32_t y0, y1;  // tensor-map rows of the boxes of output rows 0 / 1; < 0: nothing to stage
    float wy0x, wy0y, wy1x, wy1y;  // vertical weights x 2^100: (y2 - sy) of rows 0 / 1, (sy - y1) of rows 0 / 1
};
constexpr uint32_t kRecRow1 = 1u;     // output row 1 exists (odd H: the last pair has one row)
constexpr uint32_t kRecImg0 = 2u;     // output row 0 / 1 receives image data (else background)
constexpr uint32_t kRecImg1 = 4u;
constexpr uint32_t kRecNewBand = 8u;  // first item of a (plane, band) in the warp's range: the staging code re-derives its band state
constexpr uint32_t kRecOffMask = 0xffc0u;
static_assert(sizeof(ItemRec) == 32, "ItemRec layout");

// DevCrop::pad of a TMA launch: bits 0..15 = smem row bytes of this crop's box, bits 16..31 = tensor map index
__host__ __device__ __forceinline__ int32_t crop_row_bytes(const DevCrop& c) { return c.pad & 0xFFFF; }
__host__ __device__ __forceinline__ int32_t crop_map_index(const DevCrop& c) { return (c.pad >> 16) & 0xFFFF; }

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
// Tap load of the compute loop: NOT volatile, so that the compiler may hoist the loads of the next column above the
// (volatile) stores of the previous one.  Ordering against the staging pipeline is carried by data dependences: the
// address derives from a value laundered through a volatile asm after the slot's mbarrier wait, and the loaded
// value feeds a volatile store that precedes the refill of the slot.
__device__ __forceinline__ uint32_t lds32_tap(uint32_t addr) {
    uint32_t v;
    asm("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ RowTap lds_rowtap(uint32_t addr) {
    RowTap r;
    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(r.a), "=f"(r.wy0), "=f"(r.wy1), "=r"(r.pad) : "r"(addr));
    return r;
}
__device__ __forceinline__ void sts_rowtap(uint32_t addr, const RowTap& r) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(r.a), "f"(r.wy0), "f"(r.wy1), "r"(r.pad) : "memory");
}
// One lane of the (converged) warp; lets the compiler issue warp-uniform work (TMA, mbarrier) without a per-lane loop.
__device__ __forceinline__ bool elect_one_sync() {
    uint32_t pred;
    asm volatile("{\n.reg .pred p;\nelect.sync _|p, 0xffffffff;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(pred));
    return pred != 0;
}
// Value of lane 0, known to the compiler to be warp-uniform (eligible for the uniform datapath).
__device__ __forceinline__ int uniform_i(int v) { return __shfl_sync(0xffffffffu, v, 0); }
__device__ __forceinline__ uint32_t uniform_u(uint32_t v) { return __shfl_sync(0xffffffffu, v, 0); }
// Programmatic dependent launch (no-ops when the kernel was launched without the attribute).
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait_prior_grid() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// byte k of w as the float b * 2^-133 (see file header)
__device__ __forceinline__ float u8_scaled(uint32_t w, uint32_t k) {
    return __uint_as_float(__byte_perm(w, 0u, 0x4044u | (k << 8)));
}
// SaturateCast<float, uchar> in one conversion: round to nearest even, clamp to [0, 255], NaN -> 0 (what __float2uint_rn
// followed by min(., 255) gives, saturate.cuh:127-147)
__device__ __forceinline__ uint32_t f32_to_u8_sat(float v) {
    uint32_t u;
    asm("cvt.rni.sat.u8.f32 %0, %1;" : "=r"(u) : "f"(v));
    return u;
}
// halfword k of w as the float h * 2^-141: placed at mantissa bits 8..23, any 24-bit pattern X is exactly X * 2^-149
__device__ __forceinline__ float u16_scaled(uint32_t w, uint32_t k) {
    return __uint_as_float(__byte_perm(w, 0u, k ? 0x4324u : 0x4104u));
}

// (decltype(auto): a reference into the table, or -- shared launches -- the descriptor expanded from its compact form)
template <typename Table>
__device__ __forceinline__ decltype(auto) tma_crop_of(const TmaParams&, const Table& T, int z) {
    return (T.c[z]);
}
template <>
__device__ __forceinline__ decltype(auto) tma_crop_of<TmaNoTable>(const TmaParams& K, const TmaNoTable&, int z) {
    return (K.P.crops[z]);
}
template <>
__device__ __forceinline__ decltype(auto) tma_crop_of<TmaMultiTable>(const TmaParams& K, const TmaMultiTable& T, int z) {
    const DevCropC& s = T.c[z];
    DevCrop d;
    d.m.xb = s.xb;
    d.m.y0 = s.y0;
    d.w = s.w;
    d.h = s.h;
    d.pitch = s.group;
    d.fx = s.fx;
    d.fy = s.fy;
    d.bx1 = 0;
    d.by1 = 0;
    d.bx2 = K.P.W - 1;
    d.by2 = K.P.H - 1;
    d.pad = s.pad;
    return d;
}
template <typename Table>
__device__ __forceinline__ const CUtensorMap* tma_map_of(const TmaParams&, const Table& T, const DevCrop& C) {
    return &T.m[crop_map_index(C)];
}
template <>
__device__ __forceinline__ const CUtensorMap* tma_map_of<TmaNoTable>(const TmaParams& K, const TmaNoTable&, const DevCrop& C) {
    return K.maps + crop_map_index(C);
}
template <>
__device__ __forceinline__ const CUtensorMap* tma_map_of<TmaMultiTable>(const TmaParams& K, const TmaMultiTable&, const DevCrop& C) {
    return K.maps + crop_map_index(C);
}
// First float of output plane z (float layouts).
template <typename Table>
__device__ __forceinline__ float* tma_plane_base(const TmaParams& K, const Table&, int z) {
    return K.P.out.base + (long long)z * K.P.out.z_stride;
}
template <>
__device__ __forceinline__ float* tma_plane_base<TmaMultiTable>(const TmaParams& K, const TmaMultiTable& T, int z) {
    const int g = T.c[z].group;
    return T.out_base[g] + (long long)(z - T.z_first[g]) * K.P.out.z_stride;
}

// Where the staged span of a (crop, column band) starts: first in-band output column of the band, its left tap,
// and the 8-byte element coordinate of the box.  Producer and consumers must agree, so both call this.
struct BandOrigin {
    int32_t xa;       // first output column of the band that receives image data (may exceed the band: empty)
    int32_t xe;       // last such column
    int32_t c0;       // box start coordinate (8-byte elements) in the crop's tensor map
    int32_t origin;   // crop-row byte that smem byte 0 of a staged row corresponds to (= 8*c0 - xb)
};
template <int PB = 3>  // bytes per source pixel
__device__ __forceinline__ BandOrigin band_origin(const PreprocParams& P, const TmaGeom& G, const DevCrop& C, int txi) {
    BandOrigin b;
    const int tx0 = txi * (32 * G.NPB);
    b.xa = max(tx0, C.bx1);
    b.xe = min(min(tx0 + 32 * G.NPB, P.W) - 1, C.bx2);
    const AxisTap t = axis_tap(b.xa - C.bx1, C.fx);
    const int xb = C.m.xb;
    b.c0 = ((xb + PB * t.i1) >> 4) << 1;  // the box must start on a 16-byte boundary of global memory
    b.origin = 8 * b.c0 - xb;
    return b;
}

// ------------------------------------------------------------------------------------------------
// kernel
// ------------------------------------------------------------------------------------------------
// CH_GRAY: cvtColor<*2GRAY> first (registers -> one luminance, rounded to an integer like the reference's RGB2Gray<I, float>),
// then any per-channel ops on that one value; one plane is stored.
// CH_*_ALPHA: cvtColor<*2*A> in a chain on a 3-channel source (a fourth, constant channel): the chain runs on the three
// source channels as CH_FMA_DIV / CH_GENERIC, the alpha plane receives the value the host computed by running the
// constant through the same ops (TmaParams::alpha).
enum ChainKind : int { CH_GENERIC = 0, CH_FMA_DIV = 1, CH_GRAY = 2, CH_FMA_DIV_ALPHA = 3, CH_GENERIC_ALPHA = 4 };
inline bool chain_needs_image_table(int chain) { return chain >= CH_GRAY; }  // instantiated for the image-mode tables only

// x / d for both halves with 1/d = zh + zl (div_const.cpp): FMUL2 + FFMA2.
__device__ __forceinline__ float2 div_by_const2(float2 x, float zh, float zl) {
    const float2 u = __fmul2_rn(x, make_float2(zl, zl));
    return __ffma2_rn(x, make_float2(zh, zh), u);
}

// Item walk of one warp: contiguous range, decoded once and then advanced incrementally.  Items are ordered
// (plane, band, row pair); an item costs as many column groups as its band has, and the ranges are cut so that every
// warp gets the same cost (a plane of 224 columns has a 4-group and a 3-group band).
struct ItemCursor {
    int z, txi, jp, left;
    // first item whose cumulative cost reaches w (cost counted in column groups from the start of the launch)
    static __device__ __forceinline__ int item_at(const TmaGeom& G, uint32_t w) {
        const uint32_t z = fast_div(w, G.d_w_crop);
        const uint32_t r = w - z * (uint32_t)G.w_crop;
        const uint32_t in_crop = r < (uint32_t)G.w_full ? fast_div(r, G.d_NPB)
                                                          : (uint32_t)((G.tiles_x - 1) * G.HP) + fast_div(r - (uint32_t)G.w_full, G.d_np_last);
        return (int)(z * (uint32_t)G.items_per_crop + in_crop);
    }
    __device__ __forceinline__ void init(const TmaGeom& G, int g) {
        // warp g owns the cost range [g*q + min(g, r), ... + q + (g < r)) of the launch total q * n_warps + r (< 2^31)
        const uint32_t w0 = (uint32_t)g * (uint32_t)G.share_q + (uint32_t)min(g, G.share_r);
        const uint32_t w1 = w0 + (uint32_t)G.share_q + (g < G.share_r ? 1u : 0u);
        const int i0 = item_at(G, w0);
        const int i1 = g + 1 == G.grid * kWarps ? G.total_items : item_at(G, w1);
        left = i1 - i0;
        z = (int)fast_div((uint32_t)i0, G.d_items_per_crop);
        const int rem = i0 - z * G.items_per_crop;
        txi = (int)fast_div((uint32_t)rem, G.d_HP);
        jp = rem - txi * G.HP;
    }
    __device__ __forceinline__ void next(const TmaGeom& G) {
        --left;
        if (++jp == G.HP) {
            jp = 0;
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
// NC = 4 (CV_8UC4): a pixel is one aligned word -- two loads per source row, no shifts.
// S16 (CV_16SC3 / CV_16SC4): the sign bit of every halfword is flipped (sample + 32768, an unsigned halfword), converted
// like an unsigned one, and 32768 * 2^-141 = 2^-126 is subtracted again: exact, all values are multiples of 2^-141 below 2^-125.
// DEPTH = 2 (CV_16UC3 / CV_16UC4): halfword samples; 6-byte pixels start on even bytes (one funnel shift by 0 or 16 bits
// lines three words up on halfword pairs), 8-byte pixels are two aligned words.
// PERM (CH_GRAY): v[c] is the channel PRMT selector csel[c] fetches instead of channel c.
template <int NC, int DEPTH, bool S16 = false, bool PERM = false>
__device__ __forceinline__ void gather_pair(uint32_t A0, uint32_t B0, uint32_t A1, uint32_t B1, int shl, int shr, bool edge,
                                            float wx0, float wx1, float2 wy0, float2 wy1, float2 (&v)[NC],
                                            const uint32_t* csel = nullptr) {
    const float2 w00 = __fmul2_rn(make_float2(wx0, wx0), wy0), w10 = __fmul2_rn(make_float2(wx1, wx1), wy0);
    const float2 w01 = __fmul2_rn(make_float2(wx0, wx0), wy1), w11 = __fmul2_rn(make_float2(wx1, wx1), wy1);
    if constexpr (DEPTH == 2) {
        // per source row: left / right pixel as NC halfword-floats each
        float pl[4][NC], pr[4][NC];  // rows: A0, B0, A1, B1
        const uint32_t base[4] = {A0, B0, A1, B1};
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            constexpr uint32_t kFlip = S16 ? 0x80008000u : 0u;
            const uint32_t w0 = lds32_tap(base[r]) ^ kFlip, w1 = lds32_tap(base[r] + 4) ^ kFlip, w2 = lds32_tap(base[r] + 8) ^ kFlip,
                           w3 = lds32_tap(base[r] + 12) ^ kFlip;
            if constexpr (NC == 3) {
                const uint32_t L0 = __funnelshift_r(w0, w1, shl), L1 = __funnelshift_r(w1, w2, shl), L2 = __funnelshift_r(w2, w3, shl);
                pl[r][0] = u16_scaled(L0, 0), pl[r][1] = u16_scaled(L0, 1), pl[r][2] = u16_scaled(L1, 0);
                pr[r][0] = u16_scaled(L1, 1), pr[r][1] = u16_scaled(L2, 0), pr[r][2] = u16_scaled(L2, 1);
            } else {
                pl[r][0] = u16_scaled(w0, 0), pl[r][1] = u16_scaled(w0, 1), pl[r][2] = u16_scaled(w1, 0), pl[r][3] = u16_scaled(w1, 1);
                pr[r][0] = u16_scaled(w2, 0), pr[r][1] = u16_scaled(w2, 1), pr[r][2] = u16_scaled(w3, 0), pr[r][3] = u16_scaled(w3, 1);
            }
            if (edge) {  // x2_read == x1 (interpolation.cuh:72): the right tap is the left pixel again
#pragma unroll
                for (int c = 0; c < NC; ++c) pr[r][c] = pl[r][c];
            }
        }
        auto rows = [](float a, float b) {  // the sample of both rows of the pair; signed: minus the 32768 added by the flip
            float2 t = make_float2(a, b);
            if (S16) t = __fadd2_rn(t, make_float2(-1.1754943508222875e-38f, -1.1754943508222875e-38f));  // 2^-126
            return t;
        };
#pragma unroll
        for (int c = 0; c < NC; ++c) {
            float2 t = __fmul2_rn(rows(pr[0][c], pr[2][c]), w10);
            t = __ffma2_rn(rows(pl[0][c], pl[2][c]), w00, t);
            t = __ffma2_rn(rows(pl[1][c], pl[3][c]), w01, t);
            v[c] = __ffma2_rn(rows(pr[1][c], pr[3][c]), w11, t);
        }
        return;
    }
    uint32_t al0, bl0, al1, bl1, ar0, br0, ar1, br1;
    if constexpr (NC == 4) {
        al0 = lds32_tap(A0), ar0 = lds32_tap(A0 + 4);
        bl0 = lds32_tap(B0), br0 = lds32_tap(B0 + 4);
        al1 = lds32_tap(A1), ar1 = lds32_tap(A1 + 4);
        bl1 = lds32_tap(B1), br1 = lds32_tap(B1 + 4);
    } else {
        const uint32_t am0 = lds32_tap(A0 - 4), a00 = lds32_tap(A0), a01 = lds32_tap(A0 + 4);
        const uint32_t bm0 = lds32_tap(B0 - 4), b00 = lds32_tap(B0), b01 = lds32_tap(B0 + 4);
        const uint32_t am1 = lds32_tap(A1 - 4), a10 = lds32_tap(A1), a11 = lds32_tap(A1 + 4);
        const uint32_t bm1 = lds32_tap(B1 - 4), b10 = lds32_tap(B1), b11 = lds32_tap(B1 + 4);
        // left pixel in bytes 0..2 (clamped shift: 32 = the word itself), right pixel in bytes 0..2
        al0 = __funnelshift_rc(am0, a00, shl), bl0 = __funnelshift_rc(bm0, b00, shl);
        al1 = __funnelshift_rc(am1, a10, shl), bl1 = __funnelshift_rc(bm1, b10, shl);
        ar0 = __funnelshift_r(a00, a01, shr), br0 = __funnelshift_r(b00, b01, shr);
        ar1 = __funnelshift_r(a10, a11, shr), br1 = __funnelshift_r(b10, b11, shr);
    }
    if (edge) {  // x2_read == x1 (interpolation.cuh:72): the right tap is the left pixel again
        ar0 = al0;
        br0 = bl0;
        ar1 = al1;
        br1 = bl1;
    }
#pragma unroll
    for (int c = 0; c < NC; ++c) {
        auto smp = [&](uint32_t w) { return PERM ? __uint_as_float(__byte_perm(w, 0u, csel[c])) : u8_scaled(w, c); };
        float2 t = __fmul2_rn(make_float2(smp(ar0), smp(ar1)), w10);
        t = __ffma2_rn(make_float2(smp(al0), smp(al1)), w00, t);
        t = __ffma2_rn(make_float2(smp(bl0), smp(bl1)), w01, t);
        v[c] = __ffma2_rn(make_float2(smp(br0), smp(br1)), w11, t);
    }
}

// The normalised chain on a row pair (same semantics as apply_program, two values per instruction where the
// hardware has a packed form).
template <int NC>
__device__ __forceinline__ void apply_program_pair(const DevProgram& prog, float2 (&v)[NC]) {
    if (prog.round_u8) {
#pragma unroll
        for (int c = 0; c < NC; ++c) v[c] = make_float2(round_sat_kind(v[c].x, prog.round_u8), round_sat_kind(v[c].y, prog.round_u8));
    }
    for (int i = 0; i < prog.n_ops; ++i) {
        const DevOp& op = prog.ops[i];
        switch (op.kind) {
            case DOP_FMA:
#pragma unroll
                for (int c = 0; c < NC; ++c) v[c] = __ffma2_rn(v[c], make_float2(op.a[c], op.a[c]), make_float2(op.b[c], op.b[c]));
                break;
            case DOP_MUL:
#pragma unroll
                for (int c = 0; c < NC; ++c) v[c] = __fmul2_rn(v[c], make_float2(op.a[c], op.a[c]));
                break;
            case DOP_ADD:
#pragma unroll
                for (int c = 0; c < NC; ++c) v[c] = make_float2(__fadd_rn(v[c].x, op.a[c]), __fadd_rn(v[c].y, op.a[c]));
                break;
            case DOP_DIV:
#pragma unroll
                for (int c = 0; c < NC; ++c) v[c] = make_float2(__fdiv_rn(v[c].x, op.a[c]), __fdiv_rn(v[c].y, op.a[c]));
                break;
            case DOP_DIVC:
#pragma unroll
                for (int c = 0; c < NC; ++c) v[c] = div_by_const2(v[c], op.a[c], op.b[c]);
                break;
            default:
                break;
        }
    }
}

// What staging an item needs to know about its (crop, band); recomputed only when the band changes.
struct StageBand {
    int32_t c0, rb;        // box start coordinate, staged row bytes
    const CUtensorMap* map;
};

// GEN = false: the common geometry -- IGNORE_AR, every plane used, planar output.
// GEN = true : aspect-ratio bands, unused planes, packed outputs.
// PEER = true: every store goes to K.n_dest replicas of the output tensor (fast geometry only).
// MAXNP: 32-column groups per band the instantiation is compiled for.  An instantiation for planes of at most 64 columns
// (two groups: 78 registers, six CTAs per SM instead of five) was measured on the 50-crop frames and gained nothing
// (1.81 against 1.80 us per frame: that workload is bound by DRAM traffic), so only kMaxNP is built.
constexpr int kNarrowNP = 2;
constexpr int kNarrowResident = 6;
// NC: channels = bytes of the 8-bit source pixel (3: CV_8UC3, 4: CV_8UC4); registers, chain constants and planes follow it.
// U8 = true: the common geometry with packed 8-bit output (convertTo<CV_32FCn, CV_8UCn> + write<CV_8UCn>: a plain resize).
// DEPTH: bytes per source sample (1: CV_8U, 2: CV_16U / CV_16S); S16: the samples are signed.
template <typename Table, int CHAIN, bool GEN, bool PEER = false, int MAXNP = kMaxNP, int NC = 3, bool U8 = false, int DEPTH = 1, bool S16 = false>
__global__ void __launch_bounds__(kTmaThreads, MAXNP <= kNarrowNP ? kNarrowResident : kMaxResident)
preproc_tma_kernel(const __grid_constant__ TmaParams K, const __grid_constant__ Table T) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ uint64_t bar_full[kWarps * kMaxSlots];
    __shared__ ItemRec rec_table[kWarps][2 * kTapBlock];

    const PreprocParams& P = K.P;
    const TmaGeom& G = K.G;
    const int warp = uniform_i(threadIdx.x >> 5);  // warp-uniform: cursors, slot and barrier addresses derive from it
    const int lane = threadIdx.x & 31;
    const int nslots = G.slots;
    const int W = P.W, H = P.H;
    const int TW = 32 * G.NPB;

    pdl_launch_dependents();  // the next kernel of the stream may start its prologue now

    // this warp's slot ring (128-byte aligned) and its barriers
    const uint32_t slot_bytes = (uint32_t)G.slot_bytes;
    uint32_t ring = ((smem_u32(smem_raw) + 127u) & ~127u) + (uint32_t)(warp * nslots) * slot_bytes;
    uint32_t bars = smem_u32(&bar_full[warp * kMaxSlots]);
    asm volatile("" : "+r"(ring), "+r"(bars));  // keep them in registers (the compiler would re-derive them per use)
    if (lane == 0) {
        for (int s = 0; s < nslots; ++s) mbar_init(bars + 8 * s, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();

    ItemCursor cc;  // item being computed
    cc.init(G, blockIdx.x * kWarps + warp);
    const int first_item = cc.z * G.items_per_crop + cc.txi * G.HP + cc.jp;  // launch-wide index of this warp's first item
    const int n_items = cc.left;
    uint32_t recs = smem_u32(&rec_table[warp][0]);
    asm volatile("" : "+r"(recs));

    // chain constants of the specialised shape v = fma(v, ca, cb) / cd  (source-channel order)
    constexpr bool kFmaDiv = CHAIN == CH_FMA_DIV || CHAIN == CH_FMA_DIV_ALPHA;
    constexpr bool kAlpha = CHAIN == CH_FMA_DIV_ALPHA || CHAIN == CH_GENERIC_ALPHA;
    float ca[NC], cb[NC], zh[NC], zl[NC];
    if (kFmaDiv) {
#pragma unroll
        for (int c = 0; c < NC; ++c) {
            ca[c] = K.prog_img.ops[0].a[c];
            cb[c] = K.prog_img.ops[0].b[c];
            zh[c] = K.zh[c];
            zl[c] = K.zl[c];
        }
    }
    uint32_t gsel[3] = {0u, 0u, 0u};
    float gk[3] = {0.f, 0.f, 0.f};
    if (CHAIN == CH_GRAY) {
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            gsel[c] = K.gsel[c];
            gk[c] = K.gk[c];
        }
    }
    // chain(background): value of planes z >= used and of pixels outside the aspect-ratio band
    float vb[1][NC];
#pragma unroll
    for (int c = 0; c < NC; ++c) vb[0][c] = P.bg[c];
    if (GEN) apply_program<1, NC>(P.prog, vb);

    // plane offsets (floats) of the source channels
    long long oc[NC];
#pragma unroll
    for (int c = 0; c < NC; ++c) oc[c] = (long long)P.prog.dst_chan[c] * P.out.c_stride;
    const int pxs = GEN ? P.out.px_stride : 1;
    int row_step = W * pxs;  // floats between vertically adjacent pixels
    if (!GEN) asm volatile("" : "+r"(row_step));  // in a register, not re-read from the parameters per item

    // Everything below reads source images and writes the output tensor: order it after the preceding kernel
    // unless the host proved the two independent (then only completion is ordered, at the end).
    if (G.pdl_wait) pdl_wait_prior_grid();
    if (K.maps && cc.left > 0) {
        // Tensor maps that reached global memory through a host copy must be acquired for the TMA proxy before
        // their first use by this thread (once per map: the fence also drops the descriptor cache).
        const int i_last = (cc.z * G.items_per_crop + cc.txi * G.HP + cc.jp) + cc.left - 1;
        const int z_last = (int)fast_div((uint32_t)i_last, G.d_items_per_crop);
        for (int z = cc.z; z <= z_last; ++z) {
            if (GEN && z >= P.used) break;
            const CUtensorMap* map = tma_map_of<Table>(K, T, tma_crop_of<Table>(K, T, z));
            asm volatile("fence.proxy.tensormap::generic.acquire.gpu [%0], 128;" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
        }
    }

    // ---------------- item records: lane l < 16 computes row (l & 1) of the warp's item blk * 8 + (l >> 1), the even
    // lane of each pair assembles the record ----------------
    auto compute_taps = [&](int blk) {
        const int k = blk * kTapBlock + (lane >> 1);
        RowTap t;
        t.a = kRowSkip;
        t.wy0 = t.wy1 = 0.f;
        t.pad = 0;
        int rb = 0, ymap = 0;
        bool new_band = false;
        if (lane < 2 * kTapBlock && k < n_items) {
            const int idx = first_item + k;
            const int z = (int)fast_div((uint32_t)idx, G.d_items_per_crop);
            const int rem = idx - z * G.items_per_crop;
            const int txi = (int)fast_div((uint32_t)rem, G.d_HP);
            const int y = 2 * (rem - txi * G.HP) + (lane & 1);
            new_band = k == 0 || rem == txi * G.HP;
            if (y < H) {
                t.a = kRowFill;
                if (!GEN || z < P.used) {
                    const DevCrop& C = tma_crop_of<Table>(K, T, z);
                    const int tx0 = txi * TW;
                    const bool band_ok = max(tx0, C.bx1) <= min(min(tx0 + TW, W) - 1, C.bx2);
                    rb = crop_row_bytes(C);
                    ymap = C.m.y0;
                    if (band_ok && y >= C.by1 && y <= C.by2) {
                        const AxisTap v = axis_tap(y - C.by1, C.fy);
                        t.a = (uint32_t)v.i1 | ((v.i1 + 1 > C.h - 1) ? kTapClamp : 0u);
                        t.wy0 = __fmul_rn(v.w0, kWeightScale);
                        t.wy1 = __fmul_rn(v.w1, kWeightScale);
                    }
                }
            }
        }
        // the odd lane's row joins the even lane's
        RowTap u;
        u.a = __shfl_xor_sync(0xffffffffu, t.a, 1);
        u.wy0 = __shfl_xor_sync(0xffffffffu, t.wy0, 1);
        u.wy1 = __shfl_xor_sync(0xffffffffu, t.wy1, 1);
        if (lane < 2 * kTapBlock && !(lane & 1)) {
            const bool im0 = t.a < kRowFill, im1 = u.a < kRowFill;
            // staged rows of a slot: [row 0: y1, y1+1][row 1: y1, y1+1], rb bytes each; a clamped y2 re-reads y1
            const uint32_t b0 = (t.a & kTapClamp) ? 0u : (uint32_t)rb, b1 = (u.a & kTapClamp) ? 0u : (uint32_t)rb;
            const uint32_t r0 = b0 << 16, r1 = (uint32_t)(2 * rb) | ((uint32_t)(2 * rb) + b1) << 16;
            ItemRec r;
            r.a0 = (im0 ? r0 : r1) | (u.a != kRowSkip ? kRecRow1 : 0u) | (im0 ? kRecImg0 : 0u) | (im1 ? kRecImg1 : 0u) |
                   (new_band ? kRecNewBand : 0u);
            r.a1 = im1 ? r1 : r0;
            r.wy0x = im0 ? t.wy0 : u.wy0;
            r.wy1x = im0 ? t.wy1 : u.wy1;
            r.wy0y = im1 ? u.wy0 : t.wy0;
            r.wy1y = im1 ? u.wy1 : t.wy1;
            r.y0 = im0 ? ymap + (int)(t.a & (kTapClamp - 1)) : -1;
            r.y1 = im1 ? ymap + (int)(u.a & (kTapClamp - 1)) : -1;
            const uint32_t dst = recs + (uint32_t)(((blk & 1) * kTapBlock + (lane >> 1)) * sizeof(ItemRec));
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(r.a0), "r"(r.a1), "r"(r.y0), "r"(r.y1) : "memory");
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst + 16), "f"(r.wy0x), "f"(r.wy0y), "f"(r.wy1x), "f"(r.wy1y) : "memory");
        }
        __syncwarp();
    };
    // record of the warp's k-th item
    auto rec_addr = [&](int k) { return recs + (uint32_t)((k & (2 * kTapBlock - 1)) * sizeof(ItemRec)); };

    // ---------------- staging: one lane stages the (up to) two 2-row boxes of the warp's item number ks ----------------
    StageBand sb;
    sb.c0 = sb.rb = 0;
    sb.map = nullptr;
    int ks = 0;  // number of items staged so far
    auto stage_item = [&](int slot) {
        if ((ks & (kTapBlock - 1)) == 0) compute_taps(ks / kTapBlock);
        int32_t y0, y1;
        uint32_t fl, unused;
        asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(fl), "=r"(unused), "=r"(y0), "=r"(y1) : "r"(rec_addr(ks)));
        if (uniform_u(fl & kRecNewBand)) {
            const int idx = first_item + ks;
            const int iz = (int)fast_div((uint32_t)idx, G.d_items_per_crop);
            const int itx = (int)fast_div((uint32_t)(idx - iz * G.items_per_crop), G.d_HP);
            if (!GEN || iz < P.used) {
                const DevCrop& C = tma_crop_of<Table>(K, T, iz);
                const BandOrigin b = band_origin<NC * DEPTH>(P, G, C, itx);
                // all lanes computed the same values; telling the compiler so keeps the TMA issue below loop-free
                sb.c0 = uniform_i(b.c0);
                sb.rb = uniform_i(crop_row_bytes(C));
                sb.map = tma_map_of<Table>(K, T, C);
                const unsigned long long mp = reinterpret_cast<unsigned long long>(sb.map);
                sb.map = reinterpret_cast<const CUtensorMap*>(
                    ((unsigned long long)uniform_u((uint32_t)(mp >> 32)) << 32) | uniform_u((uint32_t)mp));
            }
        }
        y0 = uniform_i(y0);
        y1 = uniform_i(y1);
        if (elect_one_sync()) {
            const uint32_t sdst = ring + (uint32_t)slot * slot_bytes + kSlotHeader;
            const uint32_t full = bars + 8 * slot;
            const bool i0 = y0 >= 0, i1 = y1 >= 0;
            // the slot's previous contents were read through the generic proxy; the TMA writes through the async proxy
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
#ifdef CVGS_DIAG_SKIP_LOADS  // diagnostic build: nothing is staged (results are garbage), the barrier completes at once
            mbar_arrive_expect_tx(full, 0u);
#else
            mbar_arrive_expect_tx(full, (uint32_t)(((int)i0 + (int)i1) * 2 * sb.rb));
            if (i0) tma_load_2d(sdst, sb.map, sb.c0, y0, full);
            if (i1) tma_load_2d(sdst + 2 * sb.rb, sb.map, sb.c0, y1, full);
#endif
        }
        ++ks;
    };

    for (int s = 0; s < nslots && s < n_items; ++s) stage_item(s);

    int slot = 0, kc = 0;  // kc = number of items computed so far
    uint32_t phase = 0;
    while (cc.left > 0) {
        // ---------------- horizontal state of this lane for the band (z, txi): column p is tx0 + lane + 32 p ------
        const int z = cc.z;
        const int tx0 = cc.txi * TW;
        const int np = (min(TW, W - tx0) + 31) >> 5;
        int32_t off[MAXNP], shl[MAXNP], shr[MAXNP];
        float wxa[MAXNP], wxb[MAXNP];
        // bit p: right tap of column p clamped / column p receives image data / column p < W.  One register each,
        // tested with one LOP3 per column (separate flags were re-derived from scratch in front of every column)
        uint32_t m_edge = 0, m_img = 0, m_in = 0;
        uint32_t rb = 0;                              // staged row bytes of this crop
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
                b = band_origin<NC * DEPTH>(P, G, C, cc.txi);
                fx = C.fx;
                bx1 = C.bx1;
                wm1 = C.w - 1;
                rb = (uint32_t)crop_row_bytes(C);
            }
#pragma unroll
            for (int p = 0; p < MAXNP; ++p) {
                off[p] = shl[p] = shr[p] = 0;
                wxa[p] = wxb[p] = 0.f;
                if (p < np) {  // warp-uniform: narrow planes (64 columns: two groups) skip the unused groups
                    const int x = tx0 + lane + 32 * p;
                    const bool in_p = x < W, img_p = in_p && x >= b.xa && x <= b.xe;
                    const AxisTap t = axis_tap((img_p ? x : b.xa) - bx1, fx);
                    wxa[p] = t.w0;
                    wxb[p] = t.w1;
                    m_in |= (in_p ? 1u : 0u) << p;
                    m_img |= (img_p ? 1u : 0u) << p;
                    m_edge |= (t.i1 + 1 > wm1 ? 1u : 0u) << p;
                    const int o = NC * DEPTH * t.i1 - b.origin;
                    if (DEPTH == 2) {  // halfword samples: the word that holds the pixel's first sample, and 0 or 16 bits to shift
                        off[p] = (o >> 2) * 4;
                        shl[p] = (o & 2) * 8;
                    } else if (NC == 4) {  // word-aligned pixels (the plan requires 4-byte aligned rows)
                        off[p] = o;
                    } else {
                        off[p] = ((o + 3) >> 2) * 4;
                        shl[p] = (o & 3) ? (o & 3) * 8 : 32;
                        shr[p] = ((o + 3) & 3) * 8;
                    }
                }
            }
        }
        // this lane's first column in rows 2*jp of the three channel planes of plane z
        asm volatile("" : "+r"(m_edge), "+r"(m_img), "+r"(m_in));
        const bool full_band = tx0 + 32 * np <= W;  // every lane owns a column in each of the band's np groups
        float* sp[NC];  // this lane's first column in row 2*jp of each channel's plane
        long long rs[NC];  // floats between vertically adjacent pixels
        {
            float* const base = tma_plane_base<Table>(K, T, z) + ((long long)(tx0 + lane) * pxs + (long long)(2 * cc.jp) * row_step);
#pragma unroll
            for (int c = 0; c < NC; ++c) {
                sp[c] = base + oc[c];
                rs[c] = row_step;
            }
        }
        // 8-bit packed output (general instantiation only): byte address of this lane's first pixel in row 2*jp
        uint8_t* u8c[NC];  // per source channel: its byte in this lane's first pixel of row 2*jp
        const long long u8pitch = P.out.row_pitch;
        if ((GEN || U8) && P.out.u8) {
            uint8_t* const u8row = reinterpret_cast<uint8_t*>(P.out.base) + (long long)z * P.out.z_stride + (long long)(2 * cc.jp) * u8pitch +
                                   (long long)NC * (tx0 + lane);
#pragma unroll
            for (int c = 0; c < NC; ++c) u8c[c] = u8row + P.prog.dst_chan[c];
        }
        if (GEN && P.out.planes) {  // per-plane destinations (fk::SplitWrite): own pointer and pitch per channel
#pragma unroll
            for (int c = 0; c < NC; ++c) {
                const DevPlane pl = P.out.planes[z * NC + c];
                rs[c] = pl.pitch;
                sp[c] = pl.data + (tx0 + lane) + (long long)(2 * cc.jp) * rs[c];
            }
        }

        const int nitems = min(cc.left, G.HP - cc.jp);  // items of this band inside the warp's range
        if (!GEN) asm volatile("" : "+r"(rb));  // keep the crop's row bytes in a register (else re-read from the parameters per item)
        // The item loop is instantiated per band shape (number of column groups; ragged or not) and chosen once per
        // band: inside it the columns of a row pair are straight-line code.  Bands whose 32-column groups are all inside
        // the plane (the common case) run a branch-free body: one basic block, so the loads of the next column are
        // scheduled above the arithmetic and the stores of the previous one.  Ragged bands test the lane's column mask
        // per column.
        auto run_band = [&](auto npc_tag, auto check_tag) {
        constexpr int NPC = decltype(npc_tag)::value;
        constexpr bool CHECK = decltype(check_tag)::value;
#pragma unroll 1
        for (int it = 0; it < nitems; ++it) {
            mbar_wait(bars + 8 * slot, phase);
            const uint32_t sdata = ring + (uint32_t)slot * slot_bytes + kSlotHeader;
            uint32_t ra0, ra1;
            float2 wy0, wy1;
            {
                const uint32_t ra = rec_addr(kc);
                asm volatile("ld.shared.v2.b32 {%0, %1}, [%2];" : "=r"(ra0), "=r"(ra1) : "r"(ra));
                asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=f"(wy0.x), "=f"(wy0.y), "=f"(wy1.x), "=f"(wy1.y) : "r"(ra + 16));
            }
            const bool st1 = (ra0 & kRecRow1) != 0;
            const bool im0 = !GEN || (ra0 & kRecImg0) != 0, im1 = !GEN || (ra0 & kRecImg1) != 0;
            uint32_t aA0 = sdata + (ra0 & kRecOffMask), aB0 = sdata + (ra0 >> 16);
            uint32_t aA1 = sdata + (ra1 & 0xffffu), aB1 = sdata + (ra1 >> 16);
            // row pointers of the pair; opaque to the compiler so that they stay in registers instead of being
            // re-derived in front of every store
            float* tp[NC];
#pragma unroll
            for (int c = 0; c < NC; ++c) {
                tp[c] = sp[c] + (GEN ? rs[c] : (long long)row_step);
                asm volatile("" : "+l"(sp[c]), "+l"(tp[c]));
            }
            asm volatile("" : "+r"(aA0), "+r"(aB0), "+r"(aA1), "+r"(aB1), "+f"(wy0.x), "+f"(wy0.y), "+f"(wy1.x), "+f"(wy1.y));
            {
#pragma unroll
                for (int p = 0; p < NPC; ++p) {
                    if (!CHECK || (m_in & (1u << p))) {  // lanes past the right border of the plane skip
                        float2 v[NC];
                        if (!GEN || im0 || im1) {
                            gather_pair<NC, DEPTH, S16, CHAIN == CH_GRAY>(aA0 + off[p], aB0 + off[p], aA1 + off[p], aB1 + off[p], shl[p], shr[p],
                                        (m_edge & (1u << p)) != 0, wxa[p], wxb[p], wy0, wy1, v, gsel);
                            if (kFmaDiv) {
#pragma unroll
                                for (int c = 0; c < NC; ++c) {
                                    v[c] = __ffma2_rn(v[c], make_float2(ca[c], ca[c]), make_float2(cb[c], cb[c]));
                                    v[c] = div_by_const2(v[c], zh[c], zl[c]);
                                }
                            } else if (CHAIN == CH_GRAY) {
                                // v[0..2] arrive in operand order (gsel): FMUL on the first, FFMA on the second and third -- or, under
                                // CVGS_FP_SEPARATE, every product and sum rounded on its own; the 2^33 that undoes the tap / weight
                                // scaling is folded into the coefficients (exact)
                                float2 t;
                                if (K.gsep) {
                                    const float2 a = __fmul2_rn(v[0], make_float2(gk[0], gk[0])), b = __fmul2_rn(v[1], make_float2(gk[1], gk[1]));
                                    const float2 c2 = __fmul2_rn(v[2], make_float2(gk[2], gk[2]));
                                    t = make_float2(__fadd_rn(__fadd_rn(a.x, b.x), c2.x), __fadd_rn(__fadd_rn(a.y, b.y), c2.y));
                                } else {
                                    t = __ffma2_rn(v[2], make_float2(gk[2], gk[2]),
                                                   __ffma2_rn(v[1], make_float2(gk[1], gk[1]), __fmul2_rn(v[0], make_float2(gk[0], gk[0]))));
                                }
                                float2 g = make_float2(static_cast<float>(__float2int_rn(t.x)), static_cast<float>(__float2int_rn(t.y)));
                                if (K.gmode == 2) {  // the ops behind the conversion in canonical form
                                    g = __ffma2_rn(g, make_float2(K.ga, K.ga), make_float2(K.gb, K.gb));
                                    g = div_by_const2(g, K.gzh, K.gzl);
                                } else {
                                    for (int i = 1; i < K.prog_img.n_ops; ++i) {  // ... or one by one, on the one channel
                                        const DevOp& op = K.prog_img.ops[i];
                                        const float a = op.a[0], b = op.b[0];
                                        switch (op.kind) {
                                            case DOP_FMA: g = __ffma2_rn(g, make_float2(a, a), make_float2(b, b)); break;
                                            case DOP_MUL: g = __fmul2_rn(g, make_float2(a, a)); break;
                                            case DOP_ADD: g = make_float2(__fadd_rn(g.x, a), __fadd_rn(g.y, a)); break;
                                            case DOP_DIV: g = make_float2(__fdiv_rn(g.x, a), __fdiv_rn(g.y, a)); break;
                                            default: break;
                                        }
                                    }
                                }
                                v[0] = g;
                            } else {
                                if (G.explicit_prescale) {
#pragma unroll
                                    for (int c = 0; c < NC; ++c) v[c] = __fmul2_rn(v[c], make_float2(G.prescale, G.prescale));
                                }
                                apply_program_pair<NC>(K.prog_img, v);
                            }
                        }
                        if (GEN) {
#pragma unroll
                            for (int c = 0; c < NC; ++c) {
                                if (!(im0 && (m_img & (1u << p)))) v[c].x = vb[0][c];
                                if (!(im1 && (m_img & (1u << p)))) v[c].y = vb[0][c];
                            }
                        }
                        if ((GEN || U8) && P.out.u8) {  // SaturateCast<float, uchar> (or fk::Cast) + packed pixels, NC bytes each
                            // the byte of source channel c sits at u8c[c] (= row + dst_chan[c]); which cast is a uniform branch,
                            // not a select over both conversions
                            if (P.out.u8 == 2) {
#pragma unroll
                                for (int c = 0; c < NC; ++c) {
                                    uint8_t* ub = u8c[c] + 32 * NC * p;
                                    *ub = (uint8_t)__float2uint_rz(v[c].x);
                                    if (st1) ub[u8pitch] = (uint8_t)__float2uint_rz(v[c].y);
                                }
                            } else {
#pragma unroll
                                for (int c = 0; c < NC; ++c) {
                                    uint8_t* ub = u8c[c] + 32 * NC * p;
                                    *ub = (uint8_t)f32_to_u8_sat(v[c].x);
                                    if (st1) ub[u8pitch] = (uint8_t)f32_to_u8_sat(v[c].y);
                                }
                            }
                        } else if (PEER) {
                            const int q = 32 * p * pxs;
#pragma unroll 1
                            for (int d = 0; d < K.n_dest; ++d) {  // this GPU's tensor first, then the peers' over NVLink
                                const long long dq = K.dest_delta[d] + q;
#pragma unroll
                                for (int c = 0; c < NC; ++c) st_cs_f32(sp[c] + dq, v[c].x);
#pragma unroll
                                for (int c = 0; c < NC; ++c) st_cs_f32_if(st1, tp[c] + dq, v[c].y);
                            }
                        } else {
                            const int q = 32 * p * pxs;
#ifdef CVGS_DIAG_SKIP_STORES  // diagnostic build (scripts/diag_build.sh): stores only for a value that never occurs
                            if (v[0].x == 123456.0f)
#endif
                            {
#pragma unroll
                            for (int c = 0; c < (CHAIN == CH_GRAY ? 1 : NC); ++c) st_cs_f32(sp[c] + q, v[c].x);
#pragma unroll
                            for (int c = 0; c < (CHAIN == CH_GRAY ? 1 : NC); ++c) st_cs_f32_if(st1, tp[c] + q, v[c].y);
                            if (kAlpha) {  // the constant fourth plane
                                st_cs_f32(sp[0] + K.alpha_delta + q, K.alpha);
                                st_cs_f32_if(st1, tp[0] + K.alpha_delta + q, K.alpha);
                            }
                            }
                        }
                    }
                }
            }
#pragma unroll
            for (int c = 0; c < NC; ++c) sp[c] += 2 * (GEN ? rs[c] : (long long)row_step);
            if ((GEN || U8) && P.out.u8) {
#pragma unroll
                for (int c = 0; c < NC; ++c) u8c[c] += 2 * u8pitch;
            }

            // every lane has consumed its taps of this slot (their values fed the stores above): refill it
            __syncwarp();
            if (ks < n_items) stage_item(slot);
            ++kc;
            if (++slot == nslots) {
                slot = 0;
                phase ^= 1u;
            }
        }
        };
        using std::integral_constant;
        if (full_band) {
            if (MAXNP >= 4 && np == 4) run_band(integral_constant<int, MAXNP >= 4 ? 4 : 1>{}, integral_constant<bool, false>{});
            else if (MAXNP >= 3 && np == 3) run_band(integral_constant<int, MAXNP >= 3 ? 3 : 1>{}, integral_constant<bool, false>{});
            else if (np == 2) run_band(integral_constant<int, 2>{}, integral_constant<bool, false>{});
            else run_band(integral_constant<int, 1>{}, integral_constant<bool, false>{});
        } else {
            run_band(integral_constant<int, MAXNP>{}, integral_constant<bool, true>{});
        }
        // advance the compute cursor past the band's items
        cc.left -= nitems;
        cc.jp += nitems;
        if (cc.jp == G.HP) {
            cc.jp = 0;
            if (++cc.txi == G.tiles_x) {
                cc.txi = 0;
                ++cc.z;
            }
        }
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
inline int band_row_bytes(int TW, float fx, int pb = 3) {
    // taps of TW columns: floor((TW-1)*fx) + 2 pixels, +1 for the float rounding of the two products; pb bytes per pixel
    const double px = std::ceil(static_cast<double>(TW - 1) * static_cast<double>(fx)) + 3.0;
    const long long bytes = static_cast<long long>(px) * pb + 15 /*16-byte aligned start*/ + 4 /*w[+1] word*/;
    return static_cast<int>((bytes + 63) / 64 * 64);
}

// Host mirror of ItemCursor::init (same integer arithmetic): first item and item count of warp g.
inline void host_item_range(const TmaGeom& G, int g, long long& first, long long& count) {
    auto item_at = [&](uint32_t w) -> long long {
        const uint32_t z = w / static_cast<uint32_t>(G.w_crop);
        const uint32_t r = w - z * static_cast<uint32_t>(G.w_crop);
        const uint32_t in_crop = r < static_cast<uint32_t>(G.w_full)
                                     ? r / static_cast<uint32_t>(G.NPB)
                                     : static_cast<uint32_t>((G.tiles_x - 1) * G.HP) + (r - static_cast<uint32_t>(G.w_full)) / static_cast<uint32_t>(G.np_last);
        return static_cast<long long>(z) * G.items_per_crop + in_crop;
    };
    const uint32_t w0 = static_cast<uint32_t>(g) * static_cast<uint32_t>(G.share_q) + static_cast<uint32_t>(std::min(g, G.share_r));
    const uint32_t w1 = w0 + static_cast<uint32_t>(G.share_q) + (g < G.share_r ? 1u : 0u);
    first = item_at(w0);
    const long long last = g + 1 == G.grid * kWarps ? static_cast<long long>(G.total_items) : item_at(w1);
    count = last - first;
}

// Staged row bytes are rounded up to a few classes so that crops of one image share tensor maps.
// fine: classes every 64 bytes (launches whose maps live in the device-resident cache, where their number does not
// matter): 9 % less source traffic from DRAM on the 50-crop frames than the coarse classes, which exist so that the
// crops of a launch share the 16 maps that fit its kernel parameters.
inline int rb_class(int rb, bool fine = false) {
    if (fine) return rb <= 2048 ? std::max(128, (rb + 63) / 64 * 64) : 0;
    static const int cls[] = {128, 192, 256, 384, 512, 640, 768, 896, 1024, 1280, 1408, 1536, 1664, 1792, 1920, 2048};
    for (int c : cls)
        if (rb <= c) return c;
    return 0;
}

// cvtColor<*2GRAY> as the first op of the chain, in the geometry the CH_GRAY instantiation is built for: CV_8UC3 source,
// float interpolation, common geometry, one float plane out.
inline bool gray_program(const PreprocParams& P) {
    const DevProgram& g = P.prog;
    if (P.src_type != CVGS_8UC3 || !g.special || g.nc_out != 1 || g.nregs != 3 || g.round_u8 || g.n_ops < 1) return false;
    if ((g.ops[0].kind & 0xff) != DOP_GRAY || g.dst_chan[0] != 0) return false;
    for (int i = 1; i < g.n_ops; ++i)
        if (g.ops[i].kind != DOP_FMA && g.ops[i].kind != DOP_MUL && g.ops[i].kind != DOP_ADD && g.ops[i].kind != DOP_DIV) return false;
    return !P.band_test && P.used == P.n_planes && P.out.px_stride == 1 && !P.out.planes && !P.out.u8;
}

// Launch constants of the CH_GRAY instantiation (TmaParams::gsel ...): operand order and coefficients of the conversion
// (cvgs_device.cuh: DOP_GRAY -- bits 8..19 name the registers of x, y, z, bit 20 separate roundings, bit 21 "the FMUL is
// y * 0.587"), and the ops behind it in canonical form where the two-operation division is proven (same conditions as
// scaled_program: |a|, |b| in [2^-24, 2^24] or b zero, the luminance an integer in [0, 255]).
inline void gray_setup(const DevProgram& g, TmaParams& K) {
    const int kind = g.ops[0].kind;
    const int rx = (kind >> 8) & 3, ry = (kind >> 12) & 3, rz = (kind >> 16) & 3;
    const bool separate = (kind >> 20) & 1, y_first = (kind >> 21) & 1;
    const float kx = 0.299f * kPreScale, ky = 0.587f * kPreScale, kz = 0.114f * kPreScale;
    const int reg[3] = {y_first && !separate ? ry : rx, y_first && !separate ? rx : ry, rz};
    const float kk[3] = {y_first && !separate ? ky : kx, y_first && !separate ? kx : ky, kz};
    for (int c = 0; c < 3; ++c) {
        K.gsel[c] = 0x4044u | static_cast<uint32_t>(reg[c]) << 8;
        K.gk[c] = kk[c];
    }
    K.gsep = separate ? 1 : 0;
    K.gmode = 3;
    K.ga = 1.f;
    K.gb = -0.f;
    K.gzh = 1.f;
    K.gzl = 0.f;
    const int n = g.n_ops - 1;
    const DevOp* ops = g.ops + 1;
    const bool first_lin = n >= 1 && (ops[0].kind == DOP_MUL || ops[0].kind == DOP_FMA || ops[0].kind == DOP_ADD);
    const bool lin_only = n == 0 || (n == 1 && first_lin);
    if (!(lin_only || (n == 2 && first_lin && ops[1].kind == DOP_DIV) || (n == 1 && ops[0].kind == DOP_DIV))) return;
    const float a = !first_lin ? 1.0f : (ops[0].kind == DOP_ADD ? 1.0f : ops[0].a[0]);
    const float b = !first_lin ? -0.0f : (ops[0].kind == DOP_MUL ? -0.0f : (ops[0].kind == DOP_ADD ? ops[0].a[0] : ops[0].b[0]));
    const float d = lin_only ? 1.0f : ops[n - 1].a[0];
    const float aa = std::fabs(a), ab = std::fabs(b);
    if (!(std::isfinite(aa) && aa >= 5.9604644775390625e-08f && aa <= 16777216.0f &&
          (ab == 0.f || (std::isfinite(ab) && ab >= 5.9604644775390625e-08f && ab <= 16777216.0f))))
        return;
    const DivConst dc = div_const_prepare(d);
    const bool neg_zero_possible = ab == 0.f && std::signbit(b) && std::signbit(a);
    if (!dc.exact || !div_const_pos_zero_ok(dc) || (neg_zero_possible && !div_const_neg_zero_ok(dc))) return;
    K.gmode = 2;
    K.ga = a;
    K.gb = b;
    K.gzh = dc.zh;
    K.gzl = dc.zl;
}

// cvtColor<*2*A> (AddOpaqueAlpha) in a chain on a CV_8UC3 source, in the geometry the CH_*_ALPHA instantiations are built
// for: the program starts with the hoisted DOP_SET of register 3 (preproc_host.hpp: build_program), everything behind it is
// per-channel arithmetic, four planes of a float tensor are written.
inline bool alpha_program(const PreprocParams& P) {
    const DevProgram& g = P.prog;
    if (P.src_type != CVGS_8UC3 || !g.special || g.nc_out != 4 || g.nregs != 4 || g.n_ops < 1 || g.ops[0].kind != DOP_SET) return false;
    if (g.ops[0].b[0] != 0.f || g.ops[0].b[1] != 0.f || g.ops[0].b[2] != 0.f || g.ops[0].b[3] == 0.f) return false;
    for (int r = 0; r < 4; ++r)
        if (g.dst_chan[r] < 0) return false;
    for (int i = 1; i < g.n_ops; ++i)
        if (g.ops[i].kind != DOP_FMA && g.ops[i].kind != DOP_MUL && g.ops[i].kind != DOP_ADD && g.ops[i].kind != DOP_DIV) return false;
    return !P.band_test && P.used == P.n_planes && P.out.px_stride == 1 && !P.out.planes && !P.out.u8;
}
// The program of an alpha_program without its DOP_SET (the three source channels' chain) and the alpha register run through
// the ops on the host: the same IEEE single-precision operations the direct-gather kernel applies to it per pixel.
inline float alpha_strip(const PreprocParams& P, PreprocParams& Q) {
    Q = P;
    volatile float a = P.prog.ops[0].a[3];
    for (int i = 1; i < P.prog.n_ops; ++i) {
        const DevOp& op = P.prog.ops[i];
        const float x = a;
        switch (op.kind) {
            case DOP_FMA: a = std::fmaf(x, op.a[3], op.b[3]); break;
            case DOP_MUL: a = x * op.a[3]; break;
            case DOP_ADD: a = x + op.a[3]; break;
            default: a = x / op.a[3]; break;
        }
        Q.prog.ops[i - 1] = op;
    }
    Q.prog.n_ops = P.prog.n_ops - 1;
    Q.prog.special = 0;
    Q.prog.nc_out = 3;
    Q.prog.nregs = 3;
    return a;
}

// Second half of a launch plan, shared by the kernels built on the item / ring scheme: given the band width (G.NPB,
// G.tiles_x, G.HP, G.items_per_crop, G.total_items) and the bytes of a staging slot (G.slot_bytes), choose the ring depth,
// the CTAs per SM and the grid, and cut the items into per-warp ranges of equal cost.
inline bool tma_plan_items(TmaGeom& G, int W, int n_planes, int sm_count, int items_per_warp, int grid_div, int max_resident) {
    const int NPB = G.NPB, TW = 32 * G.NPB;
    const long long total = G.total_items;
    // ring depth: as deep as possible while kMaxResident CTAs stay resident per SM, but at least two slots
    const int smem_sm = 227 * 1024;
    // dynamic ring + static (tap tables, barriers) + the 1 KB the driver reserves per CTA
    auto cta_bytes = [&](int slots) {
        return kWarps * slots * G.slot_bytes + kRingPad + 128 + kWarps * 4 * kTapBlock * static_cast<int>(sizeof(RowTap)) + 256 + 1024;
    };
    int slots = kMaxSlots;
    while (slots > 2 && smem_sm / cta_bytes(slots) < max_resident) --slots;
    if (const char* e = std::getenv("CVGS_TMA_SLOTS")) {  // tuning override (tests / profiling)
        const int v = std::atoi(e);
        if (v >= 1 && v <= kMaxSlots) slots = v;
    }
    if (cta_bytes(slots) > smem_sm) return false;
    G.slots = slots;
    G.resident = std::min(max_resident, smem_sm / cta_bytes(slots));
    // Small launches: one item per warp spreads the work over the most SMs (shortest isolated launch); when
    // consecutive launches overlap, several items per warp amortise the staging latency and leave CTA slots free
    // for the next launch, which raises back-to-back throughput (items_per_warp > 1, cvgs_b200_set_overlap).
    const long long warps_wanted = std::max<long long>(1, (total + items_per_warp - 1) / items_per_warp);
    const long long ctas_wanted = (warps_wanted + kWarps - 1) / kWarps;
    static const int grid_mult = [] {  // tuning override (profiling): more CTAs than fit at once (waves instead of persistence)
        const char* e = std::getenv("CVGS_TMA_GRID_MULT");
        return e ? std::max(1, std::atoi(e)) : 1;
    }();
    G.grid = static_cast<int32_t>(std::min<long long>(
        ctas_wanted, std::max<long long>(1, static_cast<long long>(G.resident) * sm_count * grid_mult / std::max(1, grid_div))));
    G.np_last = (std::min(TW, W - (G.tiles_x - 1) * TW) + 31) / 32;
    const long long w_full = static_cast<long long>(G.tiles_x - 1) * G.HP * NPB;
    const long long w_crop = w_full + static_cast<long long>(G.HP) * G.np_last;
    const long long w_total = w_crop * n_planes;
    if (w_total > 0x7fffffffLL) return false;
    G.w_full = static_cast<int32_t>(w_full);
    G.w_crop = static_cast<int32_t>(w_crop);
    const long long n_warps = static_cast<long long>(G.grid) * kWarps;
    G.share_q = static_cast<int32_t>(w_total / n_warps);
    G.share_r = static_cast<int32_t>(w_total % n_warps);
    G.d_w_crop = fast_div_make(static_cast<uint32_t>(G.w_crop));
    G.d_NPB = fast_div_make(static_cast<uint32_t>(G.NPB));
    G.d_np_last = fast_div_make(static_cast<uint32_t>(G.np_last));
    G.d_items_per_crop = fast_div_make(static_cast<uint32_t>(G.items_per_crop));
    G.d_HP = fast_div_make(static_cast<uint32_t>(G.HP));
    return true;
}

// Can this launch take the TMA kernel, and with which geometry?  crops = host copies of the DevCrops.
// grid_div > 1: the launch takes only 1/grid_div of the CTA slots of the device, so that consecutive launches of a
// stream (chained by programmatic dependent launch) are co-resident and each one's ramp-up and tail overlap its neighbours.
inline bool tma_plan(const PreprocParams& P, const DevCrop* crops, int used, int n_planes, int sm_count, bool image_mode,
                     int items_per_warp, TmaGeom& G, bool need_driver = true, int grid_div = 1, int max_resident = kMaxResident) {
    if (need_driver && !encode_tiled_fn()) return false;
    // the tap extraction is written for 3-byte pixels, for 4-byte pixels that are aligned words, and for 16-bit samples
    // (6-byte pixels on even bytes, 8-byte pixels as aligned word pairs)
    if (P.src_type != CVGS_8UC3 && P.src_type != CVGS_8UC4 && P.src_type != CVGS_16UC3 && P.src_type != CVGS_16UC4 &&
        P.src_type != CVGS_16SC3 && P.src_type != CVGS_16SC4)
        return false;
    const int pb = P.src_type == CVGS_8UC3 ? 3 : (P.src_type == CVGS_8UC4 ? 4 : (P.src_type == CVGS_16UC3 || P.src_type == CVGS_16SC3 ? 6 : 8));
    if (P.out.u8 && (pb != 3 || P.prog.nc_out != 3)) return false;
    // everything but CV_8UC3: built for the common geometry only (IGNORE_AR, every plane used, planar float tensors)
    if (pb != 3 && (P.band_test || P.used != P.n_planes || P.out.px_stride != 1 || P.out.planes || P.out.u8)) return false;
    if (P.prog.special && !gray_program(P) && !alpha_program(P)) return false;  // other conversions that change the channel count: direct-gather kernel
    if (P.out.row_stride != static_cast<long long>(P.W) * P.out.px_stride) return false;  // padded packed rows: direct-gather kernel
    float fx_max = 0.f;
    for (int i = 0; i < used; ++i) {
        const DevCrop& c = crops[i];
        if (c.h > 1 && c.pitch % 16 != 0) return false;  // TMA: row stride must be a multiple of 16 bytes
        if (pb == 4 && (reinterpret_cast<uintptr_t>(c.data) & 3)) return false;  // CV_8UC4: pixels are aligned words
        if (pb == 8 && (reinterpret_cast<uintptr_t>(c.data) & 7)) return false;  // CV_16UC4: aligned word pairs
        if (!(c.fx > 0.f) || !(c.fy > 0.f) || !std::isfinite(c.fx) || !std::isfinite(c.fy)) return false;
        fx_max = std::max(fx_max, c.fx);
    }
    if (static_cast<long long>(P.W) * P.H * 4 * std::max<long long>(1, std::abs(P.out.px_stride)) > 0x3fffffffLL)
        return false;  // in-plane offsets are 32-bit in the kernel
    int NPB = std::min(kMaxNP, (P.W + 31) / 32);
    if (const char* e = std::getenv("CVGS_TMA_NPB")) {  // tuning override (profiling)
        const int v = std::atoi(e);
        if (v >= 1 && v <= kMaxNP) NPB = std::min(NPB, v);
    }
    auto need = [&](int npb) { return used > 0 ? band_row_bytes(std::min(32 * npb, P.W), fx_max, pb) : 64; };
    while (NPB > 1 && need(NPB) > kMaxBoxBytes) --NPB;
    if (need(NPB) > kMaxBoxBytes) return false;  // extreme down-scale: direct kernel
    // image mode rounds the staged row bytes up to a class (rb_class) so that crops share tensor maps
    const int rb_max = image_mode ? rb_class(need(NPB)) : need(NPB);
    const int TW = 32 * NPB;
    G.NPB = NPB;
    G.HP = (P.H + 1) / 2;
    G.tiles_x = (P.W + TW - 1) / TW;
    G.items_per_crop = G.tiles_x * G.HP;
    const long long total = static_cast<long long>(n_planes) * G.items_per_crop;
    if (total > 0x7fffffffLL) return false;
    G.total_items = static_cast<int32_t>(total);
    G.slot_bytes = kSlotHeader + 4 * rb_max;  // rb_max is a multiple of 64: slots stay 128-byte aligned
    G.explicit_prescale = 0;
    G.prescale = pb >= 6 ? kPreScale16 : kPreScale;
    G.pdl_wait = 1;
    return tma_plan_items(G, P.W, n_planes, sm_count, items_per_warp, grid_div, max_resident);
}

// The gather kernels (direct, warp, CircularTensor) run the chain from the runtime program; an IEEE division there is
// MUFU.RCP + Newton + FCHK, ~12 instructions per value.  For the canonical chain  [MUL | FMA | ADD] DIV  (or DIV alone)
// the division's numerator is fma(x, a, b) of an interpolated x in {0} U [2^-46, 2^16]: with 2^-24 <= |a|, |b|, |d| <= 2^24
// (b may be zero) a non-zero numerator has magnitude in [2^-73, 2^41] -- the range on which div_const.cpp proves the
// two-operation form exact per divisor (same argument as scaled_program below) -- and zeros are covered by the sign
// conditions.  The background value takes the same program, so it must lie in the same range.  Returns true when the
// program's DIV was replaced by DOP_DIVC.
inline bool specialize_division(DevProgram& prog, int nc, const float* bg) {
    if (prog.special) return false;
    const int n = prog.n_ops;
    DevOp* ops = prog.ops;
    const bool first_lin = n == 2 && (ops[0].kind == DOP_MUL || ops[0].kind == DOP_FMA || ops[0].kind == DOP_ADD);
    if (!((first_lin && ops[1].kind == DOP_DIV) || (n == 1 && ops[0].kind == DOP_DIV))) return false;
    const DevOp div = ops[n - 1];
    DivConst dc[4];
    for (int c = 0; c < nc; ++c) {
        const float a = n == 1 ? 1.0f : (ops[0].kind == DOP_ADD ? 1.0f : ops[0].a[c]);
        const float b = n == 1 ? -0.0f : (ops[0].kind == DOP_MUL ? -0.0f : (ops[0].kind == DOP_ADD ? ops[0].a[c] : ops[0].b[c]));
        const float aa = std::fabs(a), ab = std::fabs(b), ag = std::fabs(bg[c]);
        const bool ok = std::isfinite(aa) && aa >= 5.9604644775390625e-08f && aa <= 16777216.0f &&
                        (ab == 0.f || (std::isfinite(ab) && ab >= 5.9604644775390625e-08f && ab <= 16777216.0f)) &&
                        (ag == 0.f || (ag >= 1.4210854715202004e-14f && ag <= 65536.0f));
        if (!ok) return false;
        dc[c] = div_const_prepare(div.a[c]);
        const bool neg_zero_possible = ab == 0.f && std::signbit(b) && std::signbit(a);
        if (!dc[c].exact || !div_const_pos_zero_ok(dc[c]) || (!neg_zero_possible ? false : !div_const_neg_zero_ok(dc[c]))) return false;
    }
    DevOp& d = ops[n - 1];
    d.kind = DOP_DIVC;
    for (int c = 0; c < 4; ++c) {
        d.a[c] = c < nc ? dc[c].zh : 1.0f;
        d.b[c] = c < nc ? dc[c].zl : 0.0f;
    }
    return true;
}

// Chain for interpolated values: the 2^33 that undoes the tap/weight scaling is folded into the first op when
// that is exact for every input (MUL/FMA/DIV by a constant of moderate magnitude), else applied explicitly.
inline int scaled_program_uncached(const PreprocParams& P, TmaParams& K);
// Frame loops repeat one chain: the derived program (and the division proofs it looks up) is memoised per thread.
inline int scaled_program(const PreprocParams& P, TmaParams& K) {
    struct Memo {
        bool valid = false;
        DevProgram key;
        DevProgram prog_img;
        float zh[4], zl[4];
        int explicit_prescale, chain;
        float prescale = 0.f;
    };
    static thread_local Memo m;
    if (m.valid && m.prescale == K.G.prescale && std::memcmp(&m.key, &P.prog, sizeof(DevProgram)) == 0) {
        K.prog_img = m.prog_img;
        std::memcpy(K.zh, m.zh, sizeof m.zh);
        std::memcpy(K.zl, m.zl, sizeof m.zl);
        K.G.explicit_prescale = m.explicit_prescale;
        if (m.chain == CH_GRAY) gray_setup(P.prog, K);  // a handful of launch constants, not memoised
        return m.chain;
    }
    for (int c = 0; c < 4; ++c) K.zh[c] = K.zl[c] = 0.f;
    const int chain = scaled_program_uncached(P, K);
    m.key = P.prog;
    m.prog_img = K.prog_img;
    std::memcpy(m.zh, K.zh, sizeof m.zh);
    std::memcpy(m.zl, K.zl, sizeof m.zl);
    m.explicit_prescale = K.G.explicit_prescale;
    m.prescale = K.G.prescale;
    m.chain = chain;
    m.valid = true;
    return chain;
}
inline int scaled_program_uncached(const PreprocParams& P, TmaParams& K) {
    K.prog_img = P.prog;
    K.G.explicit_prescale = 1;
    if (gray_program(P)) {  // the conversion's coefficients carry the 2^33
        K.G.explicit_prescale = 0;
        gray_setup(P.prog, K);
        return CH_GRAY;
    }
    if (P.prog.round_u8) return CH_GENERIC;
    auto moderate = [](float a) { return a == 0.f || (std::fabs(a) > 1e-20f && std::fabs(a) < 1e20f); };
    const int nc = P.nc;
    auto all_moderate = [&](const float* a, bool nonzero) {
        for (int c = 0; c < nc; ++c)
            if (!moderate(a[c]) || (nonzero && a[c] == 0.f)) return false;
        return true;
    };
    DevOp* ops = K.prog_img.ops;
    const int n = K.prog_img.n_ops;
    // canonical shape  v = fma(v, a, b) / d : [MUL|FMA|ADD] DIV  or  DIV alone.  MUL(a) == FMA(a, -0) and
    // ADD(b) == FMA(1, b) bit for bit (x + -0 == x for every x; x * 1 is exact).
    // No op at all and a lone [MUL|FMA|ADD] are the same shape with a division by 1 (zh = 1, zl = 0: fma(x, 1, x * 0) == x bit
    // for bit, and no -0 can reach it): a plain resize and convertTo-with-scale chains then run on the specialised instantiation
    // instead of the interpreter (whose code -- every op kind in every column of every band shape -- is three times as long).
    const bool first_lin = n >= 1 && (ops[0].kind == DOP_MUL || ops[0].kind == DOP_FMA || ops[0].kind == DOP_ADD);
    const bool lin_only = n == 0 || (n == 1 && first_lin);
    if (lin_only || (n == 2 && first_lin && ops[1].kind == DOP_DIV) || (n == 1 && ops[0].kind == DOP_DIV)) {
        DevOp lin{};
        lin.kind = DOP_FMA;
        DevOp div{};
        div.kind = DOP_DIV;
        for (int c = 0; c < 4; ++c) div.a[c] = 1.0f;
        if (!lin_only) div = ops[n - 1];
        const bool has_lin = first_lin;
        for (int c = 0; c < nc; ++c) {
            lin.a[c] = !has_lin ? 1.0f : (ops[0].kind == DOP_ADD ? 1.0f : ops[0].a[c]);
            lin.b[c] = !has_lin ? -0.0f : (ops[0].kind == DOP_MUL ? -0.0f : (ops[0].kind == DOP_ADD ? ops[0].a[c] : ops[0].b[c]));
        }
        // The two-operation division is proven for normal numerators (div_const.cpp).  With 2^-24 <= |a|, |b|, |d|
        // <= 2^24 (b may be zero) a non-zero fma(x, a, b) of an interpolated x in {0} U [2^-46, 255] has magnitude
        // in [2^-73, 2^33]: every intermediate stays normal.  Zeros are handled through the signs of zh / zl.
        bool ok = true;
        DivConst dc[4];
        for (int c = 0; c < nc && ok; ++c) {
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
            for (int c = 0; c < nc; ++c) {
                lin.a[c] *= K.G.prescale;
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
        for (int c = 0; c < nc; ++c) op.a[c] *= K.G.prescale;
        K.G.explicit_prescale = 0;
    } else if (op.kind == DOP_DIV) {
        if (!all_moderate(op.a, true)) return CH_GENERIC;
        for (int c = 0; c < nc; ++c) op.a[c] /= K.G.prescale;
        K.G.explicit_prescale = 0;
    }
    return CH_GENERIC;
}

// scaled_program for every chain the kernel takes: a chain that adds an alpha channel is the three-channel chain behind
// the DOP_SET plus the constant plane.
inline int scaled_program_for(const PreprocParams& P, TmaParams& K) {
    K.alpha = 0.f;
    K.alpha_delta = 0;
    if (!alpha_program(P)) return scaled_program(P, K);
    PreprocParams Q;
    K.alpha = alpha_strip(P, Q);
    K.alpha_delta = (static_cast<long long>(P.prog.dst_chan[3]) - P.prog.dst_chan[0]) * P.out.c_stride;
    const int chain = scaled_program(Q, K);
    return chain == CH_FMA_DIV ? CH_FMA_DIV_ALPHA : CH_GENERIC_ALPHA;
}

inline int tma_encode(CUtensorMap* map, uintptr_t base16, long long row_bytes, int rows, long long pitch, int rb) {
    const cuuint64_t dim[2] = {static_cast<cuuint64_t>((row_bytes + 7) / 8), static_cast<cuuint64_t>(rows)};
    const cuuint64_t stride[1] = {rows > 1 ? static_cast<cuuint64_t>(pitch) : (dim[0] * 8 + 15) / 16 * 16};
    const cuuint32_t box[2] = {static_cast<cuuint32_t>(rb / 8), 2};
    const cuuint32_t estr[2] = {1, 1};
    static const CUtensorMapL2promotion promo = [] {  // tuning override (profiling): 0 none, 1 64 B, 2 128 B, 3 256 B
        const char* e = std::getenv("CVGS_TMA_L2PROMO");
        const int v = e ? std::atoi(e) : 2;
        return v == 0 ? CU_TENSOR_MAP_L2_PROMOTION_NONE : v == 1 ? CU_TENSOR_MAP_L2_PROMOTION_L2_64B
               : v == 3 ? CU_TENSOR_MAP_L2_PROMOTION_L2_256B : CU_TENSOR_MAP_L2_PROMOTION_L2_128B;
    }();
    const CUresult r = encode_tiled_fn()(map, CU_TENSOR_MAP_DATA_TYPE_UINT64, 2, reinterpret_cast<void*>(base16), dim, stride, box,
                                         estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, promo,
                                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(CVGS_ERR_INVALID_VALUE, "cuTensorMapEncodeTiled failed (" + std::to_string((int)r) + ")");
    return CVGS_OK;
}

// Per-crop map: nothing is known about the memory around the crop, so the map's bounds are the crop's own (reads
// never leave it by more than the 8-byte element that holds its last pixel; everything else is zero-filled).
inline int tma_prepare_crop(DevCrop& c, const TmaGeom& G, int W, int map_index, CUtensorMap* map, int pb = 3) {
    const uintptr_t addr = reinterpret_cast<uintptr_t>(c.data);
    const int mis = static_cast<int>(addr & 15);
    const int rb = band_row_bytes(std::min(32 * G.NPB, W), c.fx, pb);
    if (int rc = tma_encode(map, addr - mis, mis + static_cast<long long>(pb) * c.w, c.h, c.pitch, rb)) return rc;
    c.m.xb = mis;   // overwrites c.data (union)
    c.m.y0 = 0;
    c.pad = rb | (map_index << 16);
    return CVGS_OK;
}

// Per-image maps (cvgs_b200_preproc_launch_ex): the caller named the parent image of a crop, i.e. memory that is
// known to be readable, so one map per (image, row-bytes class) serves every crop of that image and is cached
// across launches -- camera buffers recur, their crops do not.
struct ImageMapCache {
    struct Entry {
        uintptr_t datastart = 0;
        int32_t pitch = 0, width = 0, height = 0, rb = 0;
        CUtensorMap map;
    };
    static constexpr int kEntries = 512;
    Entry e[kEntries];
    // width_bytes = pixel bytes x image width
    const CUtensorMap* get(uintptr_t datastart, int pitch, int width, int height, int rb) {
        const size_t hsh = (static_cast<size_t>(datastart >> 8) * 0x9E3779B97F4A7C15ull + static_cast<size_t>(rb) * 0xC2B2AE3D27D4EB4Full) >> 40;
        Entry& x = e[hsh % kEntries];
        if (x.datastart == datastart && x.pitch == pitch && x.width == width && x.height == height && x.rb == rb) return &x.map;
        const int mis = static_cast<int>(datastart & 15);
        if (tma_encode(&x.map, datastart - mis, mis + static_cast<long long>(width), height, pitch, rb) != CVGS_OK) {
            x.datastart = 0;
            return nullptr;
        }
        x.datastart = datastart;
        x.pitch = pitch;
        x.width = width;
        x.height = height;
        x.rb = rb;
        return &x.map;
    }
};

// Device-resident copy of the per-image maps, for launches that coalesce more frames than tensor maps fit in the
// kernel parameters.  Append-only: an entry is encoded on the host, uploaded once (flush(): one copy on an internal
// stream, waited for on the host -- a first-sight cost per (camera buffer, row-bytes class), never on a hit) and from
// then on referenced by index, so a map in use by a kernel in flight is never rewritten.  When the table is full the
// device is synchronised and the table starts over.  One cache per host thread, like ImageMapCache.
struct DevMapCache {
    static constexpr int kCap = 4096;          // entries (16-bit index in DevCrop::pad); 512 KB of device memory
    static constexpr int kBuckets = 2 * kCap;  // open addressing, power of two
    struct Key {
        uintptr_t datastart = 0;
        int32_t pitch = 0, width = 0, height = 0, rb = 0;
    };
    CUtensorMap* d = nullptr;   // device table
    CUtensorMap* h = nullptr;   // pinned host mirror
    Key* keys = nullptr;
    int32_t* bucket = nullptr;  // entry index + 1, 0 = empty
    int count = 0, uploaded = 0, device = -1;
    uint32_t generation = 0;    // bumped when the table starts over: indices handed out before are void
    cudaStream_t up = nullptr;
    int reserve(int dev) {
        if (device == dev && d) return CVGS_OK;
        release();
        CVGS_CUDA(cudaMalloc(reinterpret_cast<void**>(&d), sizeof(CUtensorMap) * kCap));
        CVGS_CUDA(cudaMallocHost(reinterpret_cast<void**>(&h), sizeof(CUtensorMap) * kCap));
        CVGS_CUDA(cudaStreamCreateWithFlags(&up, cudaStreamNonBlocking));
        keys = new Key[kCap];
        bucket = new int32_t[kBuckets]();
        count = uploaded = 0;
        device = dev;
        return CVGS_OK;
    }
    void release() {  // errors ignored: also runs at thread exit, possibly after the context is gone
        if (d) cudaFree(d);
        if (h) cudaFreeHost(h);
        if (up) cudaStreamDestroy(up);
        delete[] keys;
        delete[] bucket;
        d = h = nullptr;
        up = nullptr;
        keys = nullptr;
        bucket = nullptr;
        count = uploaded = 0;
        device = -1;
    }
    ~DevMapCache() { release(); }
    // index of the map of (image, rb), encoding it on a miss; -1 when the driver refuses the geometry
    int get(uintptr_t datastart, int pitch, int width, int height, int rb) {
        size_t b = ((static_cast<size_t>(datastart >> 8) * 0x9E3779B97F4A7C15ull + static_cast<size_t>(rb) * 0xC2B2AE3D27D4EB4Full) >> 40) &
                   (kBuckets - 1);
        for (;; b = (b + 1) & (kBuckets - 1)) {
            const int32_t e = bucket[b];
            if (e == 0) break;
            const Key& k = keys[e - 1];
            if (k.datastart == datastart && k.rb == rb && k.pitch == pitch && k.width == width && k.height == height) return e - 1;
        }
        if (count == kCap) {  // start over; nothing in flight may still read the old entries
            if (cudaDeviceSynchronize() != cudaSuccess) return -1;
            std::memset(bucket, 0, sizeof(int32_t) * kBuckets);
            count = uploaded = 0;
            ++generation;
            return get(datastart, pitch, width, height, rb);
        }
        const int mis = static_cast<int>(datastart & 15);
        if (tma_encode(&h[count], datastart - mis, mis + static_cast<long long>(width), height, pitch, rb) != CVGS_OK) return -1;
        keys[count] = Key{datastart, pitch, width, height, rb};
        bucket[b] = count + 1;
        return count++;
    }
    // make every entry handed out so far usable by kernels launched from now on (any stream)
    int flush() {
        if (uploaded == count) return CVGS_OK;
        CVGS_CUDA(cudaMemcpyAsync(d + uploaded, h + uploaded, sizeof(CUtensorMap) * static_cast<size_t>(count - uploaded),
                                  cudaMemcpyHostToDevice, up));
        CVGS_CUDA(cudaStreamSynchronize(up));
        uploaded = count;
        return CVGS_OK;
    }
};

// Locate crop c (c.data valid) inside its parent image: byte offset within a map row, map row, and the pad word that
// binds it to map `map_index` with row bytes rb.  The crop itself is not modified.
inline bool tma_place_in_image(const DevCrop& c, uintptr_t datastart, int width, int height, int rb, int map_index,
                               int32_t& xb, int32_t& y0_out, int32_t& pad, int pb = 3) {
    const uintptr_t addr = reinterpret_cast<uintptr_t>(c.data);
    if (addr < datastart || c.pitch <= 0) return false;
    const uintptr_t off = addr - datastart;
    // crops lie within 2^32 bytes of their parent in practice: 32-bit division is several times cheaper
    const unsigned long long y0 = off <= 0xffffffffull ? static_cast<uint32_t>(off) / static_cast<uint32_t>(c.pitch)
                                                       : off / static_cast<uintptr_t>(c.pitch);
    const long long xo = static_cast<long long>(off - y0 * static_cast<uintptr_t>(c.pitch));
    const long long PB = pb;
    if (xo + PB * c.w > PB * width || static_cast<long long>(y0) + c.h > height || (height > 1 && PB * width > c.pitch))
        return false;
    xb = static_cast<int32_t>((datastart & 15) + xo);
    y0_out = static_cast<int32_t>(y0);
    pad = rb | (map_index << 16);
    return true;
}

inline size_t tma_smem_bytes(const TmaGeom& G) {
    return static_cast<size_t>(kWarps) * G.slots * G.slot_bytes + kRingPad + 128;
}

template <typename Table, int CHAIN, bool GEN, bool PEER = false, int MAXNP = kMaxNP, int NC = 3, bool U8 = false, int DEPTH = 1, bool S16 = false>
inline int tma_launch_instance(const TmaParams& K, const Table& T, int device, cudaStream_t stream) {
    static thread_local size_t attr_set[64] = {};  // per device: dynamic shared memory opt-in already granted
    const size_t smem = tma_smem_bytes(K.G);
    const int slot = device & 63;
    auto kernel = preproc_tma_kernel<Table, CHAIN, GEN, PEER, MAXNP, NC, U8, DEPTH, S16>;
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

// Coalesced frame groups: only the common geometry (IGNORE_AR, every plane used, planar float output) is coalesced.
inline int tma_launch_multi(const TmaParams& K, const TmaMultiTable& T, int chain, int device, cudaStream_t stream) {
    static_assert(sizeof(TmaParams) + sizeof(TmaMultiTable) <= 32 * 1024 - 256, "kernel parameters exceed 32 KB");
    if (chain == CH_FMA_DIV) return tma_launch_instance<TmaMultiTable, CH_FMA_DIV, false>(K, T, device, stream);
    return tma_launch_instance<TmaMultiTable, CH_GENERIC, false>(K, T, device, stream);
}

// Replicated output (K.n_dest tensors): descriptors in global memory, common geometry only.
inline int tma_launch_replicated(const TmaParams& K, int chain, int device, cudaStream_t stream) {
    if (chain == CH_FMA_DIV) return tma_launch_instance<TmaNoTable, CH_FMA_DIV, false, true>(K, TmaNoTable{0}, device, stream);
    return tma_launch_instance<TmaNoTable, CH_GENERIC, false, true>(K, TmaNoTable{0}, device, stream);
}

template <typename Table>
inline int tma_launch_kernel(const TmaParams& K, const Table& T, int chain, int device, cudaStream_t stream) {
    static_assert(sizeof(TmaParams) + sizeof(Table) <= 32 * 1024, "kernel parameters exceed 32 KB");
    const PreprocParams& P = K.P;
    const bool fast = !P.band_test && P.used == P.n_planes && P.out.px_stride == 1 && !P.out.planes && !P.out.u8;
    if (P.src_type != CVGS_8UC3) {  // CV_8UC4, CV_16UC3, CV_16UC4: built for the common geometry and for the tables
                                    // batches of such frames arrive in
        if constexpr (std::is_same<Table, TmaImageTableL>::value || std::is_same<Table, TmaNoTable>::value) {
            if (!fast) return fail(CVGS_ERR_NOT_SUPPORTED, "TMA-staged kernel: this pixel type takes the common geometry only");
            const bool fd = chain == CH_FMA_DIV;
            switch (P.src_type) {
                case CVGS_8UC4:
                    return fd ? tma_launch_instance<Table, CH_FMA_DIV, false, false, kMaxNP, 4>(K, T, device, stream)
                              : tma_launch_instance<Table, CH_GENERIC, false, false, kMaxNP, 4>(K, T, device, stream);
                case CVGS_16UC3:
                    return fd ? tma_launch_instance<Table, CH_FMA_DIV, false, false, kMaxNP, 3, false, 2>(K, T, device, stream)
                              : tma_launch_instance<Table, CH_GENERIC, false, false, kMaxNP, 3, false, 2>(K, T, device, stream);
                case CVGS_16SC3:
                    return fd ? tma_launch_instance<Table, CH_FMA_DIV, false, false, kMaxNP, 3, false, 2, true>(K, T, device, stream)
                              : tma_launch_instance<Table, CH_GENERIC, false, false, kMaxNP, 3, false, 2, true>(K, T, device, stream);
                case CVGS_16SC4:
                    return fd ? tma_launch_instance<Table, CH_FMA_DIV, false, false, kMaxNP, 4, false, 2, true>(K, T, device, stream)
                              : tma_launch_instance<Table, CH_GENERIC, false, false, kMaxNP, 4, false, 2, true>(K, T, device, stream);
                default:
                    return fd ? tma_launch_instance<Table, CH_FMA_DIV, false, false, kMaxNP, 4, false, 2>(K, T, device, stream)
                              : tma_launch_instance<Table, CH_GENERIC, false, false, kMaxNP, 4, false, 2>(K, T, device, stream);
            }
        } else {
            return fail(CVGS_ERR_NOT_SUPPORTED, "TMA-staged kernel: this pixel type is not built for this descriptor table");
        }
    }
    if (chain_needs_image_table(chain)) {
        if constexpr (std::is_same<Table, TmaMultiTable>::value || std::is_same<Table, TmaParamTable>::value) {
            return fail(CVGS_ERR_NOT_SUPPORTED, "TMA-staged kernel: gray / alpha conversions are not built for this descriptor table");
        } else {
            if (chain == CH_GRAY) return tma_launch_instance<Table, CH_GRAY, false>(K, T, device, stream);
            if (!fast) return fail(CVGS_ERR_NOT_SUPPORTED, "TMA-staged kernel: alpha conversions take the common geometry only");
            return chain == CH_FMA_DIV_ALPHA ? tma_launch_instance<Table, CH_FMA_DIV_ALPHA, false>(K, T, device, stream)
                                             : tma_launch_instance<Table, CH_GENERIC_ALPHA, false>(K, T, device, stream);
        }
    }
    // packed 8-bit output of the common geometry (a plain cv::cuda::resize on CV_8UC3 is this): its own fast instantiation
    if (!P.band_test && P.used == P.n_planes && P.out.u8 && P.nc == 3 && !std::is_same<Table, TmaParamTable>::value) {
        if (chain == CH_FMA_DIV) return tma_launch_instance<Table, CH_FMA_DIV, false, false, kMaxNP, 3, true>(K, T, device, stream);
        return tma_launch_instance<Table, CH_GENERIC, false, false, kMaxNP, 3, true>(K, T, device, stream);
    }
    if (chain == CH_FMA_DIV)
        return fast ? tma_launch_instance<Table, CH_FMA_DIV, false>(K, T, device, stream)
                    : tma_launch_instance<Table, CH_FMA_DIV, true>(K, T, device, stream);
    return fast ? tma_launch_instance<Table, CH_GENERIC, false>(K, T, device, stream)
                : tma_launch_instance<Table, CH_GENERIC, true>(K, T, device, stream);
}

}  // namespace cvgs
