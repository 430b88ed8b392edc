// div_const.hpp -- see div_const.cpp.
#pragma once

namespace cvgs {

struct DivConst {
    float zh, zl;  // 1/d ~= zh + zl
    bool exact;    // RN(x*zh + RN(x*zl)) == RN(x/d) for every normal x proven by exhaustion over the mantissas
};

// RN(1/d) in float.
float correctly_rounded_reciprocal(float d);
// Cached per divisor; the first call for a new divisor takes a few milliseconds (2^23 host divisions).
DivConst div_const_prepare(float d);

// Signed zeros: x = +0 gives sign(d)*0 iff  zh > 0 || signbit(zl);  x = -0 gives -sign(d)*0 iff zh < 0 || !signbit(zl).
inline bool div_const_pos_zero_ok(const DivConst& c) { return c.zh > 0.f || __builtin_signbit(c.zl); }
inline bool div_const_neg_zero_ok(const DivConst& c) { return c.zh < 0.f || !__builtin_signbit(c.zl); }

}  // namespace cvgs
