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
