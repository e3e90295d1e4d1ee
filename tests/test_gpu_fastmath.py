"""Accuracy of the device math primitives (csrc/fastmath.cuh) against libm in extended precision, over the argument
ranges the environments produce (and beyond): covariance pivots 1e-12..1e6, angles up to 1e12 rad, pdf-ratio exponents
down to underflow.  ulp = |device - exact| / spacing(exact); `exact` is evaluated in long double (80-bit on x86)."""
import numpy as np
import pytest

from test_gpu_parity import i2c_b200  # noqa: F401

pytestmark = pytest.mark.gpu


def ulps(got, exact_ld):
    exact = np.asarray(exact_ld, dtype=np.float64)
    sp = np.spacing(np.maximum(np.abs(exact), np.finfo(float).tiny))
    return float(np.max(np.abs((got.astype(np.longdouble) - exact_ld) / sp.astype(np.longdouble))))


def test_rsqrt_rcp(i2c_b200):
    rng = np.random.default_rng(0)
    x = np.exp(rng.uniform(np.log(1e-12), np.log(1e6), 200000))
    xl = x.astype(np.longdouble)
    y, _ = i2c_b200.capi.fastmath_probe("rsqrt", x)
    assert ulps(y, 1 / np.sqrt(xl)) <= 1.5
    xs = x * rng.choice([-1.0, 1.0], x.size)
    y, _ = i2c_b200.capi.fastmath_probe("rcp", xs)
    assert ulps(y, 1 / xs.astype(np.longdouble)) <= 1.0


@pytest.mark.parametrize("fn", ["exp_neg", "exp_neg_lat"])
def test_exp_neg(i2c_b200, fn):
    rng = np.random.default_rng(1)
    x = -np.concatenate([rng.uniform(0, 1, 100000), np.exp(rng.uniform(np.log(1e-12), np.log(700.0), 100000)), [0.0]])
    y, _ = i2c_b200.capi.fastmath_probe(fn, x)
    assert ulps(y, np.exp(x.astype(np.longdouble))) <= 2.0
    # below -708 the pdf ratio is 0 for every use: the throughput flavour clamps (3e-308), the latency flavour flushes to 0
    y, _ = i2c_b200.capi.fastmath_probe(fn, np.array([-709.0, -1e4, -1e300]))
    assert np.all(y >= 0.0) and np.all(y < 1e-307)


@pytest.mark.parametrize("fn", ["sincos", "seq_sincos"])
@pytest.mark.parametrize("scale", [1.0, 10.0, 1e3, 1e5, 1e7, 1e9, 1e12])
def test_sincos(i2c_b200, fn, scale):
    """<= 2 ulp up to 1e12 rad (the 33-bit Cody-Waite split this replaces was only good to 1e5)."""
    rng = np.random.default_rng(2)
    x = rng.uniform(-scale, scale, 200000)
    x[:8] = [0.0, np.pi / 4, -np.pi / 4, np.pi / 2, np.pi, 3 * np.pi / 2, 1e-300, -1e-8][:8]
    s, c = i2c_b200.capi.fastmath_probe(fn, x)
    xl = x.astype(np.longdouble)
    # long double sin/cos reduce with a 64-bit pi: exact enough to 1e12 * 2^-64 = 5e-8 ulp-of-1 ... use the identity check too
    if scale <= 1e5:
        assert ulps(s, np.sin(xl)) <= 2.0 and ulps(c, np.cos(xl)) <= 2.0
    else:
        # libm's double routines are correctly reduced for huge arguments (Payne-Hanek): compare at 2 ulp of the larger
        # of |result| and 2^-10 (near a zero of sin the long-double reference itself is not exact)
        sr, cr = np.sin(x), np.cos(x)
        assert np.max(np.abs(s - sr) / np.spacing(np.maximum(np.abs(sr), 2.0 ** -10))) <= 3.0
        assert np.max(np.abs(c - cr) / np.spacing(np.maximum(np.abs(cr), 2.0 ** -10))) <= 3.0
    assert np.max(np.abs(s * s + c * c - 1.0)) < 1e-15


def test_sincos_beyond_fast_range(i2c_b200):
    """|x| > 1e12: the sequenced routine falls back to the library; the branch-free one stays finite and bounded."""
    x = np.array([1e13, -3e14, 1e15, 2e15])
    s, c = i2c_b200.capi.fastmath_probe("seq_sincos", x)
    assert np.max(np.abs(s - np.sin(x))) < 1e-15 and np.max(np.abs(c - np.cos(x))) < 1e-15
    s, c = i2c_b200.capi.fastmath_probe("sincos", x)
    assert np.all(np.isfinite(s)) and np.all(np.abs(s) <= 1.0 + 1e-12) and np.all(np.abs(c) <= 1.0 + 1e-12)


def test_log_accumulator(i2c_b200):
    rng = np.random.default_rng(3)
    x = np.exp(rng.uniform(np.log(1e-8), np.log(1e4), 50000))
    y, _ = i2c_b200.capi.fastmath_probe("logacc", x)
    assert np.max(np.abs(y - 3.0 * np.log(x))) < 1e-13 * np.max(np.abs(3.0 * np.log(x)))
