"""The fused kernels divide by launch constants with a 5-instruction exact sequence (Markstein) instead of the
12-instruction IEEE routine.  This test holds that sequence to __fdiv_rn for EVERY float numerator in the range
the kernels use it for (2^-60 <= |x| < 2^61, both signs: 2.03e9 values) and a set of divisors: the ones of the
BASELINE configs, awkward mantissas (all ones, just above a power of two) and random ones."""
import ctypes as C

import numpy as np
import pytest

from cvgpuspeedup_b200 import _abi

pytestmark = pytest.mark.gpu

_RNG = np.random.default_rng(99)
DIVISORS = [3.2, 0.6, 11.8, 0.229, 0.224, 0.225, 255.0, 1.0, 3.0, 7.0, 1.0 / 3.0, -5.5,
            float(np.float32(2.0) - np.float32(2.0 ** -23)),      # mantissa all ones
            float(np.float32(1.0) + np.float32(2.0 ** -23)),      # just above a power of two
            float(np.nextafter(np.float32(2.0 ** -30), np.float32(1))), float(np.float32(2.0 ** 30))] + \
           [float(np.float32(np.exp(_RNG.uniform(-18, 18)))) for _ in range(8)]


@pytest.mark.parametrize("d", DIVISORS)
def test_division_by_constant_is_ieee_exact(d):
    lib = _abi.load()
    bad = C.c_ulonglong(0)
    first = C.c_uint(0)
    rcp = C.c_float(0)
    _abi.check(lib.cvgs_b200_debug_division_sweep(C.c_float(d), C.byref(bad), C.byref(first), C.byref(rcp)))
    assert bad.value == 0, f"d={d!r} r={rcp.value!r}: {bad.value} mismatches, first numerator bits 0x{first.value:08x}"
