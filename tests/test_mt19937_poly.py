"""Host half of the MT19937 jump-ahead (csrc/mt19937.cu) against the Python restatement
(oracle/mt19937_poly.py) and against numpy's own generator.  No GPU work."""
import ctypes as C

import numpy as np
import pytest

from oracle import mt19937_poly as P
from smartstartcontinuous_b200 import _lib


def _lib_poly(J):
    lib = _lib.load()
    out = np.zeros(624, dtype=np.uint32)
    assert lib.ss_mt19937_jump_poly(C.c_uint64(J), out.ctypes.data_as(C.c_void_p)) == 0
    return int.from_bytes(out.tobytes(), "little")


def test_phi_table_matches_berlekamp_massey():
    """The exponent table compiled into the library is the characteristic polynomial Berlekamp-Massey
    finds in numpy's own output (degree 19937, 135 terms)."""
    lib = _lib.load()
    buf = (C.c_int * 256)()
    n = lib.ss_mt19937_phi_exponents(buf, 256)
    assert n == 135
    assert list(buf[:n]) == P.exponents(P.char_poly())


@pytest.mark.parametrize("J", [0, 1, 19936, 19937, 19938, 624 * 80 - 1, 2 * 131072 * 50, 2 * 1048576 * 50 + 12345,
                               (1 << 40) + 7])
def test_jump_polynomial_matches_python(J):
    assert _lib_poly(J) == P.jump_poly(J)


def test_jump_polynomial_moves_numpys_generator():
    """g = x^J mod phi applied to a RandomState key gives the key J steps later (words 1..623 and the
    top bit of word 0), checked against the sequential recurrence and numpy's own stream."""
    rs = np.random.RandomState(2024)
    rs.random_sample(1000)                                   # somewhere inside the stream
    key = rs.get_state()[1].copy()
    J = 624 * 37 - 1
    w = P.apply_jump(key, _lib_poly(J))
    x = P.raw_stream(key, J + 700)
    assert (w[1:] == x[J + 1:J + 624]).all() and (w[0] >> 31) == (x[J] >> 31)
    # the block that follows the jump target is what numpy holds after consuming 37 blocks
    pos = rs.get_state()[2]
    rs.random_sample((624 * 37 - pos) // 2 + 1)              # first draw out of block 37
    assert (rs.get_state()[1] == x[624 * 37:624 * 38]).all()
