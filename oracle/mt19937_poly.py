"""Test infrastructure (not a product path): GF(2) arithmetic behind the MT19937 jump-ahead.

numpy's legacy RandomState (the generator behind npr.uniform of NND_MB_agent.py:500-501) is MT19937:
x[k+624] = x[k+397] ^ twist(x[k], x[k+1]).  Its state sequence is linear over GF(2) with a
characteristic polynomial phi of degree 19937, so the window of 624 words J steps ahead is
    window_J[j] = XOR_{i : g_i = 1} x[i + j],      g = x^J mod phi          (Haramoto et al. 2008)
for every word j >= 1 of the window and the top bit of word 0.

  char_poly()          phi by Berlekamp-Massey over one output bit of numpy's own generator
  jump_poly(J)         x^J mod phi as a Python int (bit i = coefficient of x^i)
  apply_jump(key, g)   the window J steps ahead of `key`, by the convolution above (numpy)

csrc/mt19937.cu holds phi as a table of exponents (PHI_EXPONENTS) and computes the same polynomials
in C++; tests/test_mt19937_poly.py checks both against this file and against numpy.
"""
from __future__ import annotations

import numpy as np

N, M = 624, 397
DEG = 19937
UPPER, LOWER, MATRIX_A = 0x80000000, 0x7FFFFFFF, 0x9908B0DF


def raw_stream(key, count):
    """x[0 .. count) continuing the window `key` (x[0..624) = key): untempered words."""
    xs = np.asarray(key, dtype=np.uint32).tolist()
    for k in range(count - N):
        y = (xs[k] & UPPER) | (xs[k + 1] & LOWER)
        xs.append(xs[k + M] ^ (y >> 1) ^ (MATRIX_A if y & 1 else 0))
    return np.asarray(xs[:count], dtype=np.uint32)


def berlekamp_massey(bits):
    """Minimal polynomial of a GF(2) sequence (list of 0/1), as a Python int: connection polynomial
    C with C_0 = 1 such that sum_i C_i s[n - i] = 0; returned reversed into the characteristic form
    (bit i = coefficient of x^i, monic of degree L)."""
    n_bits = len(bits)
    C, B = 1, 1
    L, m = 0, 1
    s_int = 0       # bit (n_bits - 1 - k) = s[k]: a window of the last L + 1 values is a shift away
    for n in range(n_bits):
        s_int |= bits[n] << (n_bits - 1 - n)
    for n in range(n_bits):
        # discrepancy d = sum_{i=0..L} C_i s[n - i]; align s[n - i] with bit i
        window = (s_int >> (n_bits - 1 - n)) & ((1 << (L + 1)) - 1)
        # window bit i = s[n - i]
        d = (C & window).bit_count() & 1
        if d == 0:
            m += 1
        elif 2 * L <= n:
            T = C
            C ^= B << m
            L = n + 1 - L
            B = T
            m = 1
        else:
            C ^= B << m
            m += 1
    # characteristic polynomial: x^L * C(1/x)
    out = 0
    for i in range(L + 1):
        if (C >> i) & 1:
            out |= 1 << (L - i)
    return out, L


_PHI = None


def char_poly():
    """phi(x) of MT19937 as a Python int (degree 19937), derived from numpy's generator."""
    global _PHI
    if _PHI is None:
        rs = np.random.RandomState(12345)
        key = rs.get_state()[1]
        need = 2 * DEG + 64
        x = raw_stream(key, N + need)
        bits = [int(v) & 1 for v in x[N:N + need]]
        phi, L = berlekamp_massey(bits)
        if L != DEG:
            raise RuntimeError("Berlekamp-Massey found degree %d, expected %d" % (L, DEG))
        _PHI = phi
    return _PHI


def exponents(poly):
    return [i for i in range(poly.bit_length()) if (poly >> i) & 1]


def _reduce(a, phi_low_exps):
    mask = (1 << DEG) - 1
    while a >> DEG:
        high = a >> DEG
        a &= mask
        for e in phi_low_exps:
            a ^= high << e
    return a


def jump_poly(J, phi=None):
    """x^J mod phi (left-to-right square and multiply by x)."""
    phi = char_poly() if phi is None else phi
    low = [e for e in exponents(phi) if e < DEG]
    g = 1
    for bit in bin(J)[2:] if J > 0 else "":
        g = int(bin(g)[2:], 4)          # squaring over GF(2): spread the bits
        g = _reduce(g, low)
        if bit == "1":
            g = _reduce(g << 1, low)
    return g


def apply_jump(key, g):
    """Window J steps ahead of `key` (uint32[624]) for g = x^J mod phi: words 1..623 and the top bit of
    word 0 are those of MT19937's state after J steps."""
    x = raw_stream(key, DEG + N)
    out = np.zeros(N, dtype=np.uint32)
    for i in exponents(g):
        out ^= x[i:i + N]
    return out
