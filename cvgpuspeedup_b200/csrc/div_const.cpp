// div_const.cpp -- correctly rounded division by a launch constant in two instructions (host-side proof).
//
// The reference divides with IEEE-754 division (fk::Div, reference fkl/include/fused_kernel/algorithms/basic_ops/
// arithmetic.cuh:58-68; nvcc emits MUFU.RCP + Newton + FCHK slow path, --use_fast_math is off:
// cmake/libs/cuda/target_generation.cmake:11-12).  The divisor of cvGS::divide is a launch constant, so the kernels
// use the two-operation scheme for multiplication by an arbitrary-precision constant (Brisebarre & Muller,
// "Correctly rounded multiplication by arbitrary precision constants", IEEE TC 2008) with C = 1/d:
//
//      zh = RN(1/d)    zl = RN(1/d - zh)          (host, once per divisor)
//      u  = RN(x * zl)                             FMUL
//      q  = RN(x * zh + u)                         FFMA        q == RN(x / d) for every x  <=>  d passes the check
//
// The scheme is exact for "most" constants but not all, and the published test for a given constant is intricate.
// Both operations are homogeneous in x under scaling by powers of two and odd in x, so it is enough to try every
// mantissa once: 2^23 numerators, ~2 ms vectorised, cached per divisor.  A divisor that fails (none has been
// seen so far) sends the launch to the generic chain, which divides with __fdiv_rn.
#include "div_const.hpp"

#include <cmath>
#include <cstdint>
#include <cstring>
#include <mutex>

namespace cvgs {

float correctly_rounded_reciprocal(float d) {
    // 1.0/d in double then rounded to float can be off by one ulp in rare double-rounding cases; d * r is exact
    // in double (24 x 24 bits), so the candidate closest to 1 is picked exactly.
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

#if defined(__x86_64__) && defined(__GNUC__)
__attribute__((target("avx2,fma"))) static bool sweep_fma_hw(float d, float zh, float zl) {
    unsigned bad = 0;
    for (uint32_t m = 1u << 23; m < (1u << 24); ++m) {
        const float x = static_cast<float>(m);
        const float u = x * zl;
        const float q = __builtin_fmaf(x, zh, u);
        bad |= static_cast<unsigned>(q != x / d);
    }
    return bad == 0;
}
#endif

static bool sweep_portable(float d, float zh, float zl) {
    unsigned bad = 0;
    for (uint32_t m = 1u << 23; m < (1u << 24); ++m) {
        const float x = static_cast<float>(m);
        volatile float u = x * zl;  // a separately rounded product whatever the contraction setting
        const float q = std::fmaf(x, zh, u);
        bad |= static_cast<unsigned>(q != x / d);
    }
    return bad == 0;
}

static bool sweep(float d, float zh, float zl) {
#if defined(__x86_64__) && defined(__GNUC__)
    if (__builtin_cpu_supports("avx2") && __builtin_cpu_supports("fma")) return sweep_fma_hw(d, zh, zl);
#endif
    return sweep_portable(d, zh, zl);
}

DivConst div_const_prepare(float d) {
    struct Entry {
        uint32_t bits;
        DivConst v;
    };
    static Entry cache[128];
    static int n_cached = 0;
    static std::mutex mu;

    uint32_t bits;
    std::memcpy(&bits, &d, sizeof bits);
    {
        std::lock_guard<std::mutex> lock(mu);
        for (int i = 0; i < n_cached; ++i)
            if (cache[i].bits == bits) return cache[i].v;
    }
    DivConst r{};
    r.zh = r.zl = 0.f;
    r.exact = false;
    const float ad = std::fabs(d);
    if (std::isfinite(d) && ad >= 5.9604644775390625e-08f /*2^-24*/ && ad <= 16777216.0f /*2^24*/) {
        r.zh = correctly_rounded_reciprocal(d);
        r.zl = static_cast<float>(1.0 / static_cast<double>(d) - static_cast<double>(r.zh));
        if (r.zl == 0.f) r.zl = std::copysign(0.f, r.zh);
        // |zl| >= 2^-52 keeps x * zl normal for every numerator the kernels produce (|x| >= 2^-73)
        const bool zl_ok = r.zl == 0.f || std::fabs(r.zl) >= 2.220446049250313e-16f;
        r.exact = zl_ok && sweep(d, r.zh, r.zl);
    }
    std::lock_guard<std::mutex> lock(mu);
    if (n_cached < 128) cache[n_cached++] = Entry{bits, r};
    return r;
}

}  // namespace cvgs
