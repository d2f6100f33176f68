"""include/mas_b200/portable_math.h: accuracy against mpmath and the properties the kernels rely on."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "csrc", "portable_math_host.cpp")
LIB = os.path.join(ROOT, "tests", "_build", "libportable_math_host.so")


@pytest.fixture(scope="module")
def pm():
    hdr = os.path.join(ROOT, "include", "mas_b200", "portable_math.h")
    if not os.path.exists(LIB) or os.path.getmtime(LIB) < max(os.path.getmtime(SRC), os.path.getmtime(hdr)):
        os.makedirs(os.path.dirname(LIB), exist_ok=True)
        subprocess.check_call(["/usr/bin/g++", "-O2", "-ffp-contract=off", "-mfma", "-shared", "-fPIC", "-I" + os.path.join(ROOT, "include"), SRC,
                               "-o", LIB])
    lib = ctypes.CDLL(LIB)

    def ev(x):
        x = np.ascontiguousarray(x, dtype=np.float64)
        s, c, t = np.empty_like(x), np.empty_like(x), np.empty_like(x)
        P = ctypes.POINTER(ctypes.c_double)
        lib.pm_eval(x.ctypes.data_as(P), x.size, s.ctypes.data_as(P), c.ctypes.data_as(P), t.ctypes.data_as(P))
        return s, c, t

    return ev


def _max_ulp_error(xs, vals, fn):
    import mpmath as mp

    mp.mp.prec = 200
    worst = 0.0
    for x, v in zip(xs, vals):
        tv = fn(mp.mpf(float(x)))
        ulp = np.spacing(abs(float(tv)))
        worst = max(worst, float(abs((mp.mpf(float(v)) - tv) / mp.mpf(float(ulp)))))
    return worst


@pytest.mark.parametrize("lo,hi,n", [(-0.8, 0.8, 4000), (-10.0, 10.0, 4000), (-8e5, 8e5, 3000)])
def test_accuracy_against_mpmath(pm, lo, hi, n):
    import mpmath as mp

    xs = np.random.default_rng(0).uniform(lo, hi, n)
    s, c, t = pm(xs)
    assert _max_ulp_error(xs, s, mp.sin) < 1.0
    assert _max_ulp_error(xs, c, mp.cos) < 1.0
    assert _max_ulp_error(xs, t, mp.tan) < 2.5


def test_accuracy_near_multiples_of_half_pi(pm):
    import mpmath as mp

    rng = np.random.default_rng(1)
    k = rng.integers(-1000, 1000, 2000)
    xs = k * np.pi / 2 + rng.uniform(-1e-9, 1e-9, 2000)
    s, c, _ = pm(xs)
    assert _max_ulp_error(xs, s, mp.sin) < 1.0
    assert _max_ulp_error(xs, c, mp.cos) < 1.0


def test_special_values(pm):
    s, c, t = pm(np.array([0.0, -0.0, 1e-300, 9e5, -1e300, np.inf, np.nan]))
    assert s[0] == 0.0 and c[0] == 1.0 and t[0] == 0.0
    assert s[2] == 1e-300 and c[2] == 1.0
    assert np.all(np.isnan(s[3:])) and np.all(np.isnan(c[3:]))  # outside the documented domain


def test_symmetry_and_identity(pm):
    xs = np.random.default_rng(2).uniform(-50, 50, 5000)
    s, c, _ = pm(xs)
    sm, cm, _ = pm(-xs)
    assert np.array_equal(sm, -s) and np.array_equal(cm, c)
    assert np.max(np.abs(s * s + c * c - 1.0)) < 5e-16


def test_close_to_glibc(pm):
    """The reference uses glibc; portable results differ from it by at most one unit in the last place."""
    xs = np.random.default_rng(3).uniform(-30, 30, 20000)
    s, c, _ = pm(xs)
    assert np.max(np.abs(s - np.sin(xs)) / np.spacing(np.abs(np.sin(xs)))) <= 1.0
    assert np.max(np.abs(c - np.cos(xs)) / np.spacing(np.abs(np.cos(xs)))) <= 1.0


# ---- div_const: division by a compile-time constant through its reciprocal, correctly rounded ---------------------
DIVISORS = [2.5, 6.0, 2 * 1e-6, 1e-5 * 1e-5, 4 * 1e-5 * 1e-5, 4 * 1e-6 * 1e-6, 1e-6 * 1e-6]


def test_div_const_reciprocals_meet_the_error_condition():
    """Markstein's correction needs |RN(1/b) * b - 1| <= 2^-54 for q0 = RN(a * y) to be a faithful quotient."""
    from fractions import Fraction

    for b in DIVISORS:
        assert abs(Fraction(1.0 / b) * Fraction(b) - 1) <= Fraction(1, 2**54), b


@pytest.mark.parametrize("which", range(7))
def test_div_const_equals_division_bit_for_bit(pm, which):
    lib = ctypes.CDLL(LIB)
    lib.pm_div_const_mismatches.restype = ctypes.c_long
    rng = np.random.default_rng(100 + which)
    n = 2_000_000
    mant = rng.integers(0, 2**52, n, dtype=np.uint64)
    expo = rng.integers(1, 2047, n, dtype=np.uint64)  # every normal binade
    sign = rng.integers(0, 2, n, dtype=np.uint64)
    wide = ((sign << np.uint64(63)) | (expo << np.uint64(52)) | mant).view(np.float64)
    # mantissas of all ones / all zeros and their neighbours: the classic hard cases for reciprocal-based division
    edge_m = np.array([0, 1, 2, 2**52 - 1, 2**52 - 2, 2**51, 2**51 - 1, 2**51 + 1], dtype=np.uint64)
    edge = ((np.arange(900, 1150, dtype=np.uint64)[:, None] << np.uint64(52)) | edge_m[None, :]).reshape(-1).view(np.float64)
    near = rng.uniform(-4, 4, n)  # the magnitudes the dynamics and the finite differences see
    diff = rng.uniform(-1, 1, n) * 10.0 ** rng.uniform(-14, 2, n)
    special = np.array([0.0, -0.0, 5e-324, -5e-324, 2.2e-308, 1.7e308, -1.7e308, np.inf, -np.inf, np.nan, 1e-290, 1e290])
    P = ctypes.POINTER(ctypes.c_double)
    for arr in (wide, edge, near, diff, special):
        arr = np.ascontiguousarray(arr)
        assert lib.pm_div_const_mismatches(which, arr.ctypes.data_as(P), ctypes.c_long(arr.size)) == 0
