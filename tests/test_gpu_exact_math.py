"""The position solver's straight-line square root / reciprocal / division (kb_sqrt_u, kb_rcp_u, kb_div_u in
csrc/kb_types.cuh) must be bit-identical to the IEEE operators Box2D's x86-64 build executes (sqrtss, divss) wherever
they do not hand the input back to the plain operators: exhaustive over all 2^32 inputs for the unary ones, 2^35
pseudo-random pairs for the division."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_exact_math_matches_the_ieee_operators(native):
    bad = np.zeros(6, np.uint64)
    rc = native._fn["selftest_exact_math"](-1, 3, bad.ctypes.data_as(C.POINTER(C.c_uint64)))
    assert rc == 0, native._fn["last_error"]().decode()
    assert bad[:3].tolist() == [0, 0, 0], "sqrt / reciprocal / division mismatches: %s" % bad.tolist()
    assert bad[3] > 2 ** 34   # most inputs do take the straight-line forms (the comparison is not vacuous)
    # the kilobot-kilobot position step as the kernels run it (incl. the out-of-line exact pass) vs the plain operators
    assert bad[4] == 0, "position-pair mismatches: %d" % bad[4]
    assert 2 ** 20 < bad[5] < 2 ** 25   # the degenerate cases did go through the exact pass, the ordinary ones did not
