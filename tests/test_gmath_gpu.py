"""Device restatements of glibc powf/sinf/cosf/expf must be BIT-IDENTICAL to the host libm
(the libm the reference's rustc build links); SipHash noise must be integer-exact."""
import ctypes

import numpy as np
import pytest

from libgooey_b200 import lib, _lib

pytestmark = pytest.mark.gpu

libm = ctypes.CDLL("libm.so.6")
for f in ("powf", "sinf", "cosf", "expf"):
    getattr(libm, f).restype = ctypes.c_float
libm.powf.argtypes = [ctypes.c_float, ctypes.c_float]
for f in ("sinf", "cosf", "expf"):
    getattr(libm, f).argtypes = [ctypes.c_float]


def dev(kind, x, y=None):
    L = lib()
    x = np.ascontiguousarray(x, np.float32)
    out = np.empty_like(x)
    yp = None if y is None else np.ascontiguousarray(y, np.float32).ctypes.data_as(ctypes.c_void_p)
    L.gooey_b200_selftest_math.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint32, ctypes.c_int]
    _lib.check(L.gooey_b200_selftest_math(kind, x.ctypes.data, yp, out.ctypes.data, x.size, 0))
    return out


def bits(a):
    return np.asarray(a, np.float32).view(np.uint32)


N = 200_000


def test_powf_bit_exact():
    rng = np.random.default_rng(1)
    x = np.concatenate([rng.uniform(0, 1, N), np.full(N, 2.0), rng.uniform(0, 60, N), [0.0, 1.0, 10.0]]).astype(np.float32)
    y = np.concatenate([rng.uniform(0.1, 10, N), rng.uniform(-2, 2, N), rng.uniform(0, 3, N), [0.3, 0.0, 1.5]]).astype(np.float32)
    got = dev(0, x, y)
    want = np.array([libm.powf(float(a), float(b)) for a, b in zip(x, y)], np.float32)
    assert np.array_equal(bits(got), bits(want))


@pytest.mark.parametrize("kind,fn", [(1, "sinf"), (2, "cosf")])
def test_sincos_bit_exact(kind, fn):
    rng = np.random.default_rng(2)
    x = np.concatenate([rng.uniform(-7, 7, N), rng.uniform(-200, 200, N), rng.uniform(0, 4e5, N), rng.uniform(0, 1e9, 1000)]).astype(np.float32)
    got = dev(kind, x)
    f = getattr(libm, fn)
    want = np.array([f(float(a)) for a in x], np.float32)
    assert np.array_equal(bits(got), bits(want))


def test_expf_bit_exact():
    rng = np.random.default_rng(3)
    x = np.concatenate([rng.uniform(-20, 5, N), rng.uniform(-104, 89, N)]).astype(np.float32)
    got = dev(3, x)
    want = np.array([libm.expf(float(a)) for a in x], np.float32)
    assert np.array_equal(bits(got), bits(want))


def test_hash_noise_matches_oracle():
    import oracle_lib as O
    idx = np.arange(0, 5000, dtype=np.float32)
    got = dev(4, idx)
    want = np.array([O.lib().orc_hash_noise(int(i)) for i in idx], np.float32)
    assert np.array_equal(bits(got), bits(want))


def test_max_curve_matches_oracle():
    import oracle_lib as O
    rng = np.random.default_rng(4)
    p = rng.uniform(0, 1, 20000).astype(np.float32)
    c = rng.choice(np.array([-0.83, -0.8, -0.3, 0.8, 0.5], np.float32), 20000)
    got = dev(5, p, c)
    want = np.array([O.lib().orc_max_curve(float(a), float(b)) for a, b in zip(p, c)], np.float32)
    assert np.array_equal(bits(got), bits(want))  # powf and expm1f are both bit-exact ports


@pytest.mark.parametrize("kind,fn,lo,hi", [(6, "tanhf", -30.0, 30.0), (8, "expm1f", -30.0, 30.0), (7, "tanf", 0.0, 1.38)])
def test_fdlibm_ports_bit_exact(kind, fn, lo, hi):
    f = getattr(libm, fn)
    f.restype = ctypes.c_float
    f.argtypes = [ctypes.c_float]
    rng = np.random.default_rng(5)
    x = np.concatenate([rng.uniform(lo, hi, N), rng.uniform(max(lo, -1.0), min(hi, 1.0), N)]).astype(np.float32)
    got = dev(kind, x)
    want = np.array([f(float(a)) for a in x], np.float32)
    assert np.array_equal(bits(got), bits(want))


def test_fast_sine_of_the_additive_oscillators_stays_within_two_ulp_of_one():
    """g_sinf_fast (front end only): absolute error against f64 over the argument range a 4-minute bounce reaches."""
    rng = np.random.default_rng(6)
    x = np.concatenate([rng.uniform(-7, 7, N), rng.uniform(0, 4e5, N), rng.uniform(0, 3e7, N), [0.0, np.pi, -np.pi / 2]]).astype(np.float32)
    got = dev(9, x).astype(np.float64)
    err = np.abs(got - np.sin(x.astype(np.float64))).max()
    print(f"g_sinf_fast max abs err {err:.3e}")
    assert err <= 2.4e-7


def test_division_by_the_sample_rate_through_the_hoisted_reciprocal_is_the_ieee_quotient():
    rng = np.random.default_rng(7)
    for sr in (44100.0, 48000.0, 22050.0, 96000.0, 11025.0, 16777215.0):   # the last: the pink-noise generator's 24-bit scale (every numerator below)
        a = np.arange(1 << 24, dtype=np.float32) if sr == 16777215.0 else np.concatenate([rng.uniform(0, 1e3, N), rng.uniform(0, 3e9, N), np.exp(rng.uniform(-12, 30, N)), [0.0]]).astype(np.float32)
        b = np.full_like(a, sr)
        got = dev(10, a, b)
        assert np.array_equal(bits(got), bits(a / b)), sr
