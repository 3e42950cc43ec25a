"""The integer helpers of the resident kernel's exact grid-wide sums (arap_flow_b200/csrc/exact_limbs.cuh), compiled
for the host and checked against exact rational arithmetic: float -> four 24-bit limbs is exact inside its window and
flags everything outside it; the 4-limb total -> binary32 is rounded once, to nearest-even.  CPU only."""
import ctypes as C
import math
import os
import subprocess
from fractions import Fraction

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def el(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("el") / "libel.so")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", so, os.path.join(HERE, "exact_limbs_host.cpp")])
    L = C.CDLL(so)
    L.el_to_limbs.argtypes = [C.c_float, C.c_int, C.POINTER(C.c_int)]
    L.el_fold.argtypes = [C.POINTER(C.c_longlong), C.c_int, C.POINTER(C.c_float)]
    L.el_fold64.argtypes = [C.POINTER(C.c_longlong), C.c_int, C.POINTER(C.c_float)]
    return L


def _limbs(el, g, S):
    out = (C.c_int * 4)()
    ovf = el.el_to_limbs(C.c_float(g), S, out)
    return list(out), bool(ovf)


def _value(l, S):
    return sum(Fraction(int(v)) * Fraction(2) ** (S - 18 - 24 * k) for k, v in enumerate(l))


def test_to_limbs_exact_inside_the_window(el):
    rng = np.random.default_rng(0)
    for S in (-100, -37, 0, 5, 41, 100):
        # magnitudes from 2^(S-66) (ulp >= 2^(S-90)) up to just below 2^(S+6)
        ex = rng.integers(S - 66, S + 6, 4000)
        g = (np.ldexp(rng.uniform(1.0, 2.0, 4000), ex) * rng.choice([-1.0, 1.0], 4000)).astype(np.float32)
        for x in g:
            if not (abs(float(x)) < 2.0 ** (S + 6)):
                continue
            l, ovf = _limbs(el, float(x), S)
            assert not ovf and all(abs(v) < (1 << 24) for v in l)
            assert _value(l, S) == Fraction(float(x)), (float(x), S, l)
    # zero, negative zero
    for z in (0.0, -0.0):
        assert _limbs(el, z, 3) == ([0, 0, 0, 0], False)


def test_to_limbs_rounds_below_the_lsb_and_flags_above_the_window(el):
    S = 10
    lsb = Fraction(2) ** (S - 90)
    rng = np.random.default_rng(1)
    ex = rng.integers(S - 120, S - 66, 3000)
    g = (np.ldexp(rng.uniform(1.0, 2.0, 3000), ex) * rng.choice([-1.0, 1.0], 3000)).astype(np.float32)
    for x in g:
        l, ovf = _limbs(el, float(x), S)
        assert not ovf
        assert abs(_value(l, S) - Fraction(float(x))) <= lsb / 2      # nearest multiple of the LSB
    for x in (2.0 ** (S + 6), -2.0 ** (S + 6), 3.0 * 2.0 ** (S + 20), float("inf"), float("-inf"), float("nan")):
        l, ovf = _limbs(el, x, S)
        assert ovf and l == [0, 0, 0, 0]
    # largest value inside the window
    x = float(np.nextafter(np.float32(2.0 ** (S + 6)), np.float32(0)))
    l, ovf = _limbs(el, x, S)
    assert not ovf and _value(l, S) == Fraction(x)
    # subnormal terms against a small scale
    for x in (1e-45, -3e-45, 1.1754942e-38):
        l, ovf = _limbs(el, x, -100)
        assert not ovf and abs(_value(l, -100) - Fraction(float(np.float32(x)))) <= Fraction(2) ** (-190) / 2


def _rn_even(T, e_unit):
    """exact integer T * 2^e_unit -> binary32, round to nearest even (normal range)"""
    if T == 0:
        return 0.0
    a, nb = abs(T), abs(T).bit_length()
    sh = max(nb - 24, 0)
    q, rem = a >> sh, a & ((1 << sh) - 1)
    if sh and (rem > (1 << (sh - 1)) or (rem == (1 << (sh - 1)) and (q & 1))):
        q += 1
    return math.copysign(math.ldexp(q, sh + e_unit), T)


def test_fold_rounds_once_to_nearest_even(el):
    rng = np.random.default_rng(2)
    out = C.c_float()

    def fold(L, e_unit):
        arr = (C.c_longlong * 4)(*L)
        rc = el.el_fold(arr, e_unit, C.byref(out))
        return rc, float(out.value)

    cases = []
    for _ in range(6000):
        n = int(rng.integers(1, 5))
        L = [0, 0, 0, 0]
        for k in range(4 - n, 4):
            L[k] = int(rng.integers(-(1 << 44), 1 << 44))
        if rng.random() < 0.3:          # massive cancellation between limbs
            L[0] = int(rng.integers(-3, 4)); L[1] = -L[0] * (1 << 24) + int(rng.integers(-5, 6))
        cases.append(L)
    # ties: exactly half an ulp above an even / odd 24-bit value, and a sticky bit far below
    cases += [[0, 0, (1 << 24) | 0, 1 << 23], [0, 0, (1 << 24) | 1, 1 << 23], [0, 1 << 24, 0, 1], [1, 0, 0, -1], [0, 0, 0, 1],
              [0, 0, 0, -(1 << 24) - 1], [(1 << 44), -(1 << 44), (1 << 44), -(1 << 44)]]
    for L in cases:
        T = (L[0] << 72) + (L[1] << 48) + (L[2] << 24) + L[3]
        for e_unit in (-190, -133, -90, -40, 0, 10):
            rc, got = fold(L, e_unit)
            nb = abs(T).bit_length()
            if T != 0 and (nb + e_unit < -120 or nb + e_unit > 120):
                assert rc == 1          # leaves it to the binary64 route
                continue
            assert rc == (2 if T == 0 else 0)
            want = float(np.float32(_rn_even(T, e_unit)))
            assert got == want and math.copysign(1.0, got) == math.copysign(1.0, want if T != 0 else 0.0), (L, e_unit, got, want)


def test_fold64_fast_path_agrees_or_declines(el):
    """The 64-bit decode either declines (preconditions) or returns the correctly rounded result; on totals shaped like the
    kernel's (leading bit inside the two top limbs, signed limb sums of every sign combination) it must not decline."""
    rng = np.random.default_rng(3)
    out = C.c_float()
    taken = 0
    cases = []
    for _ in range(20000):
        L = [int(rng.integers(-(1 << 22), 1 << 22)), int(rng.integers(-(1 << 39), 1 << 39)),
             int(rng.integers(-(1 << 39), 1 << 39)), int(rng.integers(-(1 << 39), 1 << 39))]
        r = rng.random()
        if r < 0.15:
            L[0] = int(rng.integers(-3, 4))
        elif r < 0.25:
            L[0] = 0; L[1] = int(rng.integers(-(1 << 30), 1 << 30))
        elif r < 0.30:
            L[2] = 0; L[3] = 0                                  # exact ties possible
        elif r < 0.35:
            L[2] = -(1 << 24) * int(rng.integers(0, 3)); L[3] = int(rng.integers(-2, 3))
        cases.append(L)
    cases += [[0, 1 << 24, 0, 0], [0, (1 << 24) | 1, 0, 0], [0, (1 << 25) | 1, 0, 0], [0, (1 << 25) | 1, 0, 1], [0, (1 << 25) | 1, 0, -1],
              [0, (1 << 25) | 3, 0, 0], [0, -((1 << 25) | 1), 0, 0], [0, -((1 << 25) | 1), 0, 1], [0, -((1 << 25) | 3), -1, (1 << 24)],
              [1, -(1 << 24), 0, 0], [1 << 36, 0, 0, 1], [0, 0, 0, 0]]
    for L in cases:
        T = (L[0] << 72) + (L[1] << 48) + (L[2] << 24) + L[3]
        for e_unit in (-150, -90, -30, 10):
            arr = (C.c_longlong * 4)(*L)
            rc = el.el_fold64(arr, e_unit, C.byref(out))
            if rc == 1:
                continue
            taken += 1
            assert rc == (2 if T == 0 else 0), (L, rc)
            want = float(np.float32(_rn_even(T, e_unit)))
            assert float(out.value) == want, (L, e_unit, float(out.value), want)
            # kernel-shaped totals must take the fast path
    assert taken > 0.7 * len(cases) * 4
    arr = (C.c_longlong * 4)(1 << 14, 123456, -98765, 4242)
    assert el.el_fold64(arr, -60, C.byref(out)) == 0
