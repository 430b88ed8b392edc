// preproc_yuv_tma.cuh -- NV12 / NV21 frames through the TMA-staged item / ring scheme of preproc_tma.cuh.
//
// Replaces, for batches of decoder frames in the common geometry (IGNORE_AR, every plane used, planar float tensor),
// the instantiation the reference builds from fk::Resize<INTER_LINEAR>::build(fk::fuse(Read<ReadYUV<NV12>>,
// Unary<ConvertYUVToRGB<NV12, range, primaries, false, float3>>), size) + chain + TensorSplit (reference
// fkl/.../image_processing/color_conversion.cuh:235-362, tests/resize/test_fused_resize.cu:73-76,141-143): every one of
// the four bilinear taps is a source pixel converted to float RGB (per channel FMUL(y * m0), FFMA(u, m1, .),
// FFMA(v, m2, .) after y - 16 [bt601], u - 128, v - 128), the interpolation runs on those.
//
// Same scheme as preproc_tma_kernel: an item is a pair of output rows x a band of up to 128 columns; every warp owns a
// contiguous range of items and feeds itself through a private ring of shared-memory slots filled by cp.async.bulk.tensor
// (per output row one 2-row box of the luma plane and one 2-row box of the interleaved chroma plane, each through its
// own tensor map); the two rows of a pair ride in the halves of packed FP32 instructions.  The u8 -> f32 conversion is
// the same one-PRMT trick (byte at mantissa bits 16..23 = b * 2^-133); here the compensating 2^100 sits in the colour
// matrix (the offsets 16 / 128 are subtracted exactly in the scaled domain, the first product of every channel brings the
// value back into the normal range before anything is rounded) and the remaining 2^33 in the first op of the chain, so
// every rounding is the reference's.
// DEPTH = 2: P010 / P210 frames (16-bit words, the 10-bit sample in the high bits; P210 has a chroma row per luma row).
// The reference shifts the samples down by 6, converts in the 10-bit range and multiplies the float RGB by 64; here the
// low six bits are masked off and the halfword lands at mantissa bits 8..23 (= sample * 2^-135), the matrix carries 2^108
// (2^100, the 64 and the 2^2 between the two placements) -- the same values at the same 2^-33 scale, every rounding in
// the normal range.
// PACKED: Y210 frames ({Y0, U, Y1, V} 16-bit words per pixel pair, one plane, one tensor map): a tap is the 8-byte group of
// its pixel pair, fetched with one 64-bit shared-memory load; its three samples are halfwords 0 / 2 (Y), 1 (U) and 3 (V) of
// the group, each placed with one two-source PRMT and cleaned (neighbour bytes and the low six bits) with one LOP3.
// The aspect-ratio / partial-batch / packed-output forms stay with the direct kernel.
#pragma once
#include "preproc_tma.cuh"

namespace cvgs {

struct __align__(16) DevYuv {  // one frame of a launch
    int32_t xbL, xbC;    // byte offset of the luma / chroma plane's first byte within a row of its tensor map (base & 15)
    int32_t w, h;        // frame size in pixels (even)
    float fx, fy;        // src_conv_factors
    int32_t rbL, rbC;    // staged row bytes (box widths) of the two maps, multiples of 64
    int32_t mapL, mapC;  // indices into the device-resident map table
    int32_t pad0, pad1;
};
static_assert(sizeof(DevYuv) == 48, "DevYuv layout");

struct YuvParams {
    PreprocParams P;
    DevProgram prog_img;       // chain for interpolated values (2^33 folded into its first op)
    float zh[4], zl[4];        // CH_FMA_DIV
    TmaGeom G;
    const CUtensorMap* maps;   // device-resident table (DevMapCache)
    const DevYuv* frames;      // device table, one entry per plane
    float m[9];                // YCbCr -> RGB matrix x 2^100
    float yoff, coff;          // luma / chroma offsets x 2^-133
    uint32_t selU1, selV1;     // PRMT selectors of U / V of the first chroma pair of a staged word (NV12: bytes 0, 1; NV21: 1, 0;
                               // 16-bit: halfwords 0, 1)
    int32_t csh;               // chroma rows: luma row >> csh (1 for 4:2:0, 0 for 4:2:2)
};

// (Y, U, V) of one tap for both rows of the pair -> float RGB at scale 2^-33
__device__ __forceinline__ void yuv_to_rgb2(const YuvParams& K, float2 y, float2 u, float2 v, float2 (&rgb)[3]) {
    const float2 yy = __fadd2_rn(y, make_float2(-K.yoff, -K.yoff));
    const float2 uu = __fadd2_rn(u, make_float2(-K.coff, -K.coff));
    const float2 vv = __fadd2_rn(v, make_float2(-K.coff, -K.coff));
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        float2 t = __fmul2_rn(yy, make_float2(K.m[3 * r], K.m[3 * r]));
        t = __ffma2_rn(uu, make_float2(K.m[3 * r + 1], K.m[3 * r + 1]), t);
        rgb[r] = __ffma2_rn(vv, make_float2(K.m[3 * r + 2], K.m[3 * r + 2]), t);
    }
}

__device__ __forceinline__ uint2 lds64_tap(uint32_t a) {
    uint2 v;
    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(a));
    return v;
}

template <int CHAIN, int DEPTH = 1, bool PACKED = false>
__global__ void __launch_bounds__(kTmaThreads, kMaxResident)
preproc_yuv_tma_kernel(const __grid_constant__ YuvParams K) {
    constexpr int B = PACKED ? 4 : DEPTH;  // bytes per luma sample step (packed: a pixel pair is 8 bytes)
    extern __shared__ uint8_t smem_raw[];
    __shared__ uint64_t bar_full[kWarps * kMaxSlots];

    const PreprocParams& P = K.P;
    const TmaGeom& G = K.G;
    const int warp = uniform_i(threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    const int nslots = G.slots;
    const int W = P.W, H = P.H;
    const int TW = 32 * G.NPB;

    pdl_launch_dependents();

    const uint32_t slot_bytes = (uint32_t)G.slot_bytes;
    uint32_t ring = ((smem_u32(smem_raw) + 127u) & ~127u) + (uint32_t)(warp * nslots) * slot_bytes;
    uint32_t bars = smem_u32(&bar_full[warp * kMaxSlots]);
    asm volatile("" : "+r"(ring), "+r"(bars));
    if (lane == 0) {
        for (int s = 0; s < nslots; ++s) mbar_init(bars + 8 * s, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();

    ItemCursor cc;  // item being computed
    cc.init(G, blockIdx.x * kWarps + warp);
    ItemCursor ic = cc;  // item being staged (nslots ahead)

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
    long long oc[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) oc[c] = (long long)P.prog.dst_chan[c] * P.out.c_stride;
    int row_step = W;
    asm volatile("" : "+r"(row_step));

    pdl_wait_prior_grid();  // plain stream order (the descriptor copy sits in front of this kernel anyway)
    if (cc.left > 0) {  // tensor maps in global memory: acquire them for the TMA proxy (see preproc_tma_kernel)
        const int i_last = (cc.z * G.items_per_crop + cc.txi * G.HP + cc.jp) + cc.left - 1;
        const int z_last = (int)fast_div((uint32_t)i_last, G.d_items_per_crop);
        for (int z = cc.z; z <= z_last; ++z) {
            const DevYuv& F = K.frames[z];
            asm volatile("fence.proxy.tensormap::generic.acquire.gpu [%0], 128;" ::"l"(reinterpret_cast<uint64_t>(K.maps + F.mapL)) : "memory");
            asm volatile("fence.proxy.tensormap::generic.acquire.gpu [%0], 128;" ::"l"(reinterpret_cast<uint64_t>(K.maps + F.mapC)) : "memory");
        }
    }

    // ---------------- staging: per output row one luma box (rows y1, y1 + 1) and one chroma box (rows y1 >> 1, + 1) ----------
    struct {
        int z, txi;
        int32_t c0L, c0C, rbL, rbC;
        float fy;
        const CUtensorMap *mapL, *mapC;
    } sb;
    sb.z = sb.txi = -1;
    sb.c0L = sb.c0C = sb.rbL = sb.rbC = 0;
    sb.fy = 1.f;
    sb.mapL = sb.mapC = nullptr;
    auto stage_item = [&](int slot) {
        if (ic.z != sb.z || ic.txi != sb.txi) {
            sb.z = ic.z;
            sb.txi = ic.txi;
            const DevYuv& F = K.frames[ic.z];
            const AxisTap tb = axis_tap(ic.txi * TW, F.fx);
            sb.c0L = uniform_i(((F.xbL + (PACKED ? 8 * (tb.i1 >> 1) : B * tb.i1)) >> 4) << 1);
            sb.c0C = uniform_i(((F.xbC + 2 * B * (tb.i1 >> 1)) >> 4) << 1);
            sb.rbL = uniform_i(F.rbL);
            sb.rbC = uniform_i(F.rbC);
            sb.fy = __uint_as_float(uniform_u(__float_as_uint(F.fy)));
            const unsigned long long mpL = reinterpret_cast<unsigned long long>(K.maps + F.mapL);
            const unsigned long long mpC = reinterpret_cast<unsigned long long>(K.maps + F.mapC);
            sb.mapL = reinterpret_cast<const CUtensorMap*>(((unsigned long long)uniform_u((uint32_t)(mpL >> 32)) << 32) | uniform_u((uint32_t)mpL));
            sb.mapC = reinterpret_cast<const CUtensorMap*>(((unsigned long long)uniform_u((uint32_t)(mpC >> 32)) << 32) | uniform_u((uint32_t)mpC));
        }
        const int y = 2 * ic.jp;
        const int y1a = uniform_i(axis_tap(y, sb.fy).i1);
        const int y1b = uniform_i(axis_tap(y + 1, sb.fy).i1);
        const bool two = y + 1 < H;
        if (elect_one_sync()) {
            const uint32_t rs = (uint32_t)(2 * sb.rbL + (PACKED ? 0 : 2 * sb.rbC));
            const uint32_t sdst = ring + (uint32_t)slot * slot_bytes + kSlotHeader;
            const uint32_t full = bars + 8 * slot;
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            mbar_arrive_expect_tx(full, two ? 2u * rs : rs);
            tma_load_2d(sdst, sb.mapL, sb.c0L, y1a, full);
            if (!PACKED) tma_load_2d(sdst + 2 * sb.rbL, sb.mapC, sb.c0C, y1a >> K.csh, full);
            if (two) {
                tma_load_2d(sdst + rs, sb.mapL, sb.c0L, y1b, full);
                if (!PACKED) tma_load_2d(sdst + rs + 2 * sb.rbL, sb.mapC, sb.c0C, y1b >> K.csh, full);
            }
        }
        ic.next(G);
    };

    for (int s = 0; s < nslots && ic.left > 0; ++s) stage_item(s);

    int slot = 0;
    uint32_t phase = 0;
    while (cc.left > 0) {
        // ---------------- horizontal state of this lane for the band (z, txi): column p is tx0 + lane + 32 p ------
        const int z = cc.z;
        const int tx0 = cc.txi * TW;
        const int np = (min(TW, W - tx0) + 31) >> 5;
        const DevYuv& F = K.frames[z];
        const int rbL = F.rbL, rbC = F.rbC, hm1 = F.h - 1;
        const float fy = F.fy;
        int32_t offL[kMaxNP], offC[kMaxNP];
        uint32_t cfg[kMaxNP], selY2[kMaxNP], selU2[kMaxNP];
        float wxa[kMaxNP], wxb[kMaxNP];
        uint32_t m_in = 0;
        {
            const AxisTap tb = axis_tap(tx0, F.fx);
            const int originL = 8 * (((F.xbL + (PACKED ? 8 * (tb.i1 >> 1) : B * tb.i1)) >> 4) << 1) - F.xbL;  // luma-row byte smem byte 0 stands for
            const int originC = 8 * (((F.xbC + 2 * B * (tb.i1 >> 1)) >> 4) << 1) - F.xbC;  // same for the chroma row
            const int wm1 = F.w - 1;
            const uint32_t kU = (K.selU1 >> 8) & 7u;  // byte of U within a pair
#pragma unroll
            for (int p = 0; p < kMaxNP; ++p) {
                offL[p] = offC[p] = 0;
                cfg[p] = selY2[p] = selU2[p] = 0;
                wxa[p] = wxb[p] = 0.f;
                if (p < np) {
                    const int x = tx0 + lane + 32 * p;
                    const bool in_p = x < W;
                    const AxisTap t = axis_tap(in_p ? x : tx0, F.fx);
                    wxa[p] = t.w0;
                    wxb[p] = t.w1;
                    m_in |= (in_p ? 1u : 0u) << p;
                    const int x2 = min(t.i1 + 1, wm1);  // interpolation.cuh:72: the right tap is clamped to the last pixel
                    const int oL = B * t.i1 - originL, oC = 2 * B * (t.i1 >> 1) - originC;
                    offL[p] = (oL >> 2) * 4;
                    offC[p] = (oC >> 2) * 4;
                    cfg[p] = (uint32_t)((oL & 3) * 8) | (uint32_t)((oC & 3) * 8) << 8;  // funnel shifts: luma in bits 0..4, chroma 8..12
                    if (PACKED) {  // byte offsets of the two taps' groups; halfword 0 or 2 of a group is the pixel's Y
                        offL[p] = 8 * (t.i1 >> 1) - originL;
                        offC[p] = 8 * (x2 >> 1) - originL;
                        selY2[p] = (t.i1 & 1) ? 0x0540u : 0x0100u;  // left tap:  bytes (4, 5) or (0, 1) of the group -> bytes 1, 2
                        selU2[p] = (x2 & 1) ? 0x0540u : 0x0100u;    // right tap
                    } else if (DEPTH == 1) {
                        selY2[p] = 0x4044u | (uint32_t)(x2 - t.i1) << 8;
                        selU2[p] = 0x4044u | (uint32_t)(2 * ((x2 >> 1) - (t.i1 >> 1)) + (int)kU) << 8;
                    } else {  // halfword 0 / 1 of the lined-up luma word; which of two chroma words holds the right tap's pair
                        selY2[p] = x2 != t.i1 ? 0x4324u : 0x4104u;
                        selU2[p] = (x2 >> 1) != (t.i1 >> 1) ? 0x7654u : 0x3210u;
                    }
                }
            }
        }
        asm volatile("" : "+r"(m_in));
        const bool full_band = tx0 + 32 * np <= W;
        float* sp[3];
        {
            float* const base = P.out.base + ((long long)z * P.out.z_stride + (long long)(tx0 + lane) + (long long)(2 * cc.jp) * row_step);
#pragma unroll
            for (int c = 0; c < 3; ++c) sp[c] = base + oc[c];
        }
        const int nitems = min(cc.left, G.HP - cc.jp);
        auto run_band = [&](auto npc_tag, auto check_tag) {
            constexpr int NPC = decltype(npc_tag)::value;
            constexpr bool CHECK = decltype(check_tag)::value;
#pragma unroll 1
            for (int it = 0; it < nitems; ++it) {
                // vertical taps of the pair (warp-uniform arithmetic)
                const int y = 2 * (cc.jp + it);
                const bool st1 = y + 1 < H;
                const AxisTap ta = axis_tap(y, fy), tb2 = axis_tap(st1 ? y + 1 : y, fy);
                const float2 wy0 = make_float2(ta.w0, tb2.w0), wy1 = make_float2(ta.w1, tb2.w1);
                const int y2a = min(ta.i1 + 1, hm1), y2b = min(tb2.i1 + 1, hm1);
                const uint32_t rs = (uint32_t)(2 * rbL + (PACKED ? 0 : 2 * rbC));
                const uint32_t sdata = ring + (uint32_t)slot * slot_bytes + kSlotHeader;
                const uint32_t r1 = st1 ? rs : 0u;  // a missing second row borrows the first one's taps (not stored)
                // luma rows y1 / y2 and chroma rows (y1 >> 1) / (y2 >> 1) of both output rows
                uint32_t LA0 = sdata, LB0 = sdata + (y2a != ta.i1 ? (uint32_t)rbL : 0u);
                const int csh = K.csh;
                uint32_t CA0 = sdata + 2 * (uint32_t)rbL, CB0 = CA0 + ((y2a >> csh) != (ta.i1 >> csh) ? (uint32_t)rbC : 0u);
                uint32_t LA1 = sdata + r1, LB1 = LA1 + (y2b != tb2.i1 ? (uint32_t)rbL : 0u);
                uint32_t CA1 = LA1 + 2 * (uint32_t)rbL, CB1 = CA1 + ((y2b >> csh) != (tb2.i1 >> csh) ? (uint32_t)rbC : 0u);
                float* tp[3];
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    tp[c] = sp[c] + (long long)row_step;
                    asm volatile("" : "+l"(sp[c]), "+l"(tp[c]));
                }
                mbar_wait(bars + 8 * slot, phase);
                asm volatile("" : "+r"(LA0), "+r"(LB0), "+r"(CA0), "+r"(CB0), "+r"(LA1), "+r"(LB1), "+r"(CA1), "+r"(CB1));
#pragma unroll
                for (int p = 0; p < NPC; ++p) {
                    if (!CHECK || (m_in & (1u << p))) {
                        const int sL = (int)(cfg[p] & 31u), sC = (int)(cfg[p] >> 8);
                        auto lineup = [&](uint32_t a, int sh) { return __funnelshift_r(lds32_tap(a), lds32_tap(a + 4), sh); };
                        auto px = [](uint32_t w0, uint32_t w1, uint32_t sel) {
                            return make_float2(__uint_as_float(__byte_perm(w0, 0u, sel)), __uint_as_float(__byte_perm(w1, 0u, sel)));
                        };
                        float2 p00[3], p10[3], p01[3], p11[3];  // taps (x1, y1), (x2, y1), (x1, y2), (x2, y2) as RGB x 2^-33
                        if (PACKED) {
                            // one 8-byte group per tap and row: {Y0 U | Y1 V}; sample -> bytes 1, 2 of a float word, the rest cleared
                            constexpr uint32_t kKeep = 0x00ffc000u;
                            auto tap = [&](uint32_t rowA, uint32_t rowB, int32_t off, uint32_t selY, float2 (&rgb)[3]) {
                                const uint2 ga = lds64_tap(rowA + off), gb = lds64_tap(rowB + off);
                                auto smp = [&](uint32_t sel) {
                                    return make_float2(__uint_as_float(__byte_perm(ga.x, ga.y, sel) & kKeep),
                                                       __uint_as_float(__byte_perm(gb.x, gb.y, sel) & kKeep));
                                };
                                yuv_to_rgb2(K, smp(selY), smp(0x0320u), smp(0x0760u), rgb);  // U: bytes 2, 3; V: bytes 6, 7
                            };
                            tap(LA0, LA1, offL[p], selY2[p], p00);
                            tap(LA0, LA1, offC[p], selU2[p], p10);
                            tap(LB0, LB1, offL[p], selY2[p], p01);
                            tap(LB0, LB1, offC[p], selU2[p], p11);
                            (void)sL;
                            (void)sC;
                        } else if (DEPTH == 1) {
                            const uint32_t selV1 = K.selV1, selU1 = K.selU1, selV2 = selU2[p] ^ 0x100u;
                            // staged words lined up on the left tap: luma [Y(x1) Y(x1+1) ..], chroma [pair(x1 >> 1) pair(+1)]
                            const uint32_t la0 = lineup(LA0 + offL[p], sL), lb0 = lineup(LB0 + offL[p], sL);
                            const uint32_t la1 = lineup(LA1 + offL[p], sL), lb1 = lineup(LB1 + offL[p], sL);
                            const uint32_t cA0 = lineup(CA0 + offC[p], sC), cB0 = lineup(CB0 + offC[p], sC);
                            const uint32_t cA1 = lineup(CA1 + offC[p], sC), cB1 = lineup(CB1 + offC[p], sC);
                            yuv_to_rgb2(K, px(la0, la1, 0x4044u), px(cA0, cA1, selU1), px(cA0, cA1, selV1), p00);
                            yuv_to_rgb2(K, px(la0, la1, selY2[p]), px(cA0, cA1, selU2[p]), px(cA0, cA1, selV2), p10);
                            yuv_to_rgb2(K, px(lb0, lb1, 0x4044u), px(cB0, cB1, selU1), px(cB0, cB1, selV1), p01);
                            yuv_to_rgb2(K, px(lb0, lb1, selY2[p]), px(cB0, cB1, selU2[p]), px(cB0, cB1, selV2), p11);
                        } else {
                            // luma: [Y(x1) Y(x1+1)] halfwords lined up (the shift is 0 or 16); chroma: pair(x1 >> 1) is one
                            // aligned word, the right tap's pair that word or the next; the low six bits of every sample go
                            constexpr uint32_t kTen = 0xffc0ffc0u;
                            const uint32_t selV1 = K.selV1, selU1 = K.selU1, selC = selU2[p];
                            const uint32_t la0 = lineup(LA0 + offL[p], sL) & kTen, lb0 = lineup(LB0 + offL[p], sL) & kTen;
                            const uint32_t la1 = lineup(LA1 + offL[p], sL) & kTen, lb1 = lineup(LB1 + offL[p], sL) & kTen;
                            auto pairs = [&](uint32_t a, uint32_t& left, uint32_t& right) {
                                const uint32_t w0 = lds32_tap(a), w1 = lds32_tap(a + 4);
                                left = w0 & kTen;
                                right = __byte_perm(w0, w1, selC) & kTen;
                            };
                            uint32_t cA0, cA0r, cB0, cB0r, cA1, cA1r, cB1, cB1r;
                            pairs(CA0 + offC[p], cA0, cA0r);
                            pairs(CB0 + offC[p], cB0, cB0r);
                            pairs(CA1 + offC[p], cA1, cA1r);
                            pairs(CB1 + offC[p], cB1, cB1r);
                            (void)sC;
                            yuv_to_rgb2(K, px(la0, la1, 0x4104u), px(cA0, cA1, selU1), px(cA0, cA1, selV1), p00);
                            yuv_to_rgb2(K, px(la0, la1, selY2[p]), px(cA0r, cA1r, selU1), px(cA0r, cA1r, selV1), p10);
                            yuv_to_rgb2(K, px(lb0, lb1, 0x4104u), px(cB0, cB1, selU1), px(cB0, cB1, selV1), p01);
                            yuv_to_rgb2(K, px(lb0, lb1, selY2[p]), px(cB0r, cB1r, selU1), px(cB0r, cB1r, selV1), p11);
                        }
                        const float2 wxa2 = make_float2(wxa[p], wxa[p]), wxb2 = make_float2(wxb[p], wxb[p]);
                        const float2 w00 = __fmul2_rn(wxa2, wy0), w10 = __fmul2_rn(wxb2, wy0);
                        const float2 w01 = __fmul2_rn(wxa2, wy1), w11 = __fmul2_rn(wxb2, wy1);
                        float2 v[3];
#pragma unroll
                        for (int c = 0; c < 3; ++c) {
                            float2 t = __fmul2_rn(p10[c], w10);
                            t = __ffma2_rn(p00[c], w00, t);
                            t = __ffma2_rn(p01[c], w01, t);
                            v[c] = __ffma2_rn(p11[c], w11, t);
                        }
                        if (CHAIN == CH_FMA_DIV) {
#pragma unroll
                            for (int c = 0; c < 3; ++c) {
                                v[c] = __ffma2_rn(v[c], make_float2(ca[c], ca[c]), make_float2(cb[c], cb[c]));
                                v[c] = div_by_const2(v[c], zh[c], zl[c]);
                            }
                        } else {
                            if (G.explicit_prescale) {
#pragma unroll
                                for (int c = 0; c < 3; ++c) v[c] = __fmul2_rn(v[c], make_float2(G.prescale, G.prescale));
                            }
                            apply_program_pair<3>(K.prog_img, v);
                        }
                        const int q = 32 * p;
#pragma unroll
                        for (int c = 0; c < 3; ++c) st_cs_f32(sp[c] + q, v[c].x);
#pragma unroll
                        for (int c = 0; c < 3; ++c) st_cs_f32_if(st1, tp[c] + q, v[c].y);
                    }
                }
#pragma unroll
                for (int c = 0; c < 3; ++c) sp[c] += 2 * (long long)row_step;
                __syncwarp();
                if (ic.left > 0) stage_item(slot);
                if (++slot == nslots) {
                    slot = 0;
                    phase ^= 1u;
                }
            }
        };
        using std::integral_constant;
        if (full_band) {
            if (np == 4) run_band(integral_constant<int, 4>{}, integral_constant<bool, false>{});
            else if (np == 3) run_band(integral_constant<int, 3>{}, integral_constant<bool, false>{});
            else if (np == 2) run_band(integral_constant<int, 2>{}, integral_constant<bool, false>{});
            else run_band(integral_constant<int, 1>{}, integral_constant<bool, false>{});
        } else {
            run_band(integral_constant<int, kMaxNP>{}, integral_constant<bool, true>{});
        }
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
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
// Staged row bytes of the two planes for a band of TW columns at scale fx: luma 1 byte per pixel; the chroma row spans
// the same bytes (2 bytes per pixel pair) plus one pair on either side.
inline int yuv_rb_luma(int TW, float fx, int depth = 1) { return band_row_bytes(TW, fx, depth); }
inline int yuv_rb_chroma(int TW, float fx, int depth = 1) { return band_row_bytes(TW, fx, depth) + 64; }
inline int yuv_depth_of(int src_type) { return src_type == CVGS_P010 || src_type == CVGS_P210 ? 2 : 1; }
// Y210: 4 bytes per pixel; a tap's 8-byte group can start one pixel before it and end one pixel after it
inline int yuv_rb_packed(int TW, float fx) { return band_row_bytes(TW, fx, 4) + 64; }

// Can a batch of NV12 / NV21 frames take this kernel, and with which geometry?
inline bool yuv_tma_plan(const PreprocParams& P, const DevCrop* crops, int used, int n_planes, int sm_count, TmaGeom& G) {
    if (!encode_tiled_fn()) return false;
    if (P.src_type != CVGS_NV12 && P.src_type != CVGS_NV21 && P.src_type != CVGS_P010 && P.src_type != CVGS_P210 && P.src_type != CVGS_Y210)
        return false;
    const bool packed = P.src_type == CVGS_Y210;
    const int depth = packed ? 4 : yuv_depth_of(P.src_type);  // bytes per pixel of the (luma) plane
    if (P.band_test || P.used != P.n_planes || used != n_planes || P.out.px_stride != 1 || P.out.planes || P.out.u8 || P.prog.special) return false;
    float fx_max = 0.f;
    for (int i = 0; i < used; ++i) {
        const DevCrop& c = crops[i];
        if ((c.w & 1) || (c.h & 1) || c.w < 2 || c.h < 2 || c.pitch % 16 != 0 || c.pitch < depth * c.w) return false;
        if (depth == 2 && (reinterpret_cast<uintptr_t>(c.data) & 3)) return false;  // chroma pairs are aligned words
        if (packed && (reinterpret_cast<uintptr_t>(c.data) & 7)) return false;      // pixel pairs are aligned 8-byte groups
        if (!(c.fx > 0.f) || !(c.fy > 0.f) || !std::isfinite(c.fx) || !std::isfinite(c.fy)) return false;
        fx_max = std::max(fx_max, c.fx);
    }
    if (static_cast<long long>(P.W) * P.H * 4 > 0x3fffffffLL) return false;
    int NPB = std::min(kMaxNP, (P.W + 31) / 32);
    auto need = [&](int npb) { return yuv_rb_chroma(std::min(32 * npb, P.W), fx_max, depth); };
    while (NPB > 1 && need(NPB) > kMaxBoxBytes) --NPB;
    if (need(NPB) > kMaxBoxBytes) return false;
    const int TW = 32 * NPB;
    std::memset(&G, 0, sizeof G);
    G.NPB = NPB;
    G.HP = (P.H + 1) / 2;
    G.tiles_x = (P.W + TW - 1) / TW;
    G.items_per_crop = G.tiles_x * G.HP;
    const long long total = static_cast<long long>(n_planes) * G.items_per_crop;
    if (total > 0x7fffffffLL) return false;
    G.total_items = static_cast<int32_t>(total);
    G.slot_bytes = kSlotHeader + (packed ? 4 * yuv_rb_packed(std::min(TW, P.W), fx_max)
                                         : 2 * (2 * yuv_rb_luma(std::min(TW, P.W), fx_max, depth) + 2 * need(NPB)));
    G.explicit_prescale = 0;
    G.prescale = kPreScale;
    G.pdl_wait = 1;
    return tma_plan_items(G, P.W, n_planes, sm_count, 1, 1, kMaxResident);
}

template <int CHAIN, int DEPTH = 1, bool PACKED = false>
inline int yuv_launch_instance(const YuvParams& K, int device, cudaStream_t stream) {
    static thread_local size_t attr_set[64] = {};
    const size_t smem = tma_smem_bytes(K.G);
    const int slot = device & 63;
    auto kernel = preproc_yuv_tma_kernel<CHAIN, DEPTH, PACKED>;
    if (smem > attr_set[slot]) {
        const size_t want = std::max<size_t>(smem, 112 * 1024);
        CVGS_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(want)));
        attr_set[slot] = want;
    }
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
    CVGS_CUDA(cudaLaunchKernelEx(&cfg, kernel, K));
    count_launch();
    return CVGS_OK;
}

}  // namespace cvgs
