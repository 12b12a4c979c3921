"""The position solver's straight-line square root / reciprocal / division (kb_sqrt_u, kb_rcp_u, kb_div_u in
csrc/kb_types.cuh) must be bit-identical to the IEEE operators Box2D's x86-64 build executes (sqrtss, divss) wherever
they do not hand the input back to the plain operators: exhaustive over all 2^32 inputs for the unary ones, 2^35
pseudo-random pairs for the division."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_exact_math_matches_the_ieee_operators(native):
    bad = np.zeros(4, np.uint64)
    rc = native._fn["selftest_exact_math"](-1, 3, bad.ctypes.data_as(C.POINTER(C.c_uint64)))
    assert rc == 0, native._fn["last_error"]().decode()
    assert bad[:3].tolist() == [0, 0, 0], "sqrt / reciprocal / division mismatches: %s" % bad.tolist()
    assert bad[3] > 2 ** 34   # most inputs do take the straight-line forms (the comparison is not vacuous)
