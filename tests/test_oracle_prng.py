"""Pins the oracle's PRNG layer: Random123 Threefry-2x32-20 known-answer vectors,
jax.random.split / key layouts for both jax_threefry_partitionable settings (App. B), the
uniform->normal map, and erf_inv against scipy / mpmath.  Also checks that the product's
host-side key management (diffrax_b200.random) is the same function."""
import numpy as np
import pytest
import scipy.special as sp

import oracle
import diffrax_b200 as dfx

KATS = [((0, 0), (0, 0), (0x6b200159, 0x99ba4efe)),
        ((0xffffffff, 0xffffffff), (0xffffffff, 0xffffffff), (0x1cb996fc, 0xbb002be7)),
        ((0x13198a2e, 0x03707344), (0x243f6a88, 0x85a308d3), (0xc4923a9c, 0x483df7a0))]


@pytest.mark.parametrize("key,ctr,want", KATS)
def test_threefry_known_answers(key, ctr, want):
    assert oracle.threefry2x32(*key, *ctr) == want
    a, b = dfx.random.threefry2x32(key[0], key[1], ctr[0], ctr[1])
    assert (int(a), int(b)) == want


def test_key_words():
    assert list(oracle.prng_key(0)) == [0, 0]
    assert list(oracle.prng_key(42)) == [0, 42]
    assert list(oracle.prng_key((7 << 32) | 9)) == [7, 9]
    assert np.array_equal(dfx.random.key(12345), oracle.prng_key(12345))


@pytest.mark.parametrize("part", [True, False])
@pytest.mark.parametrize("num", [1, 2, 3, 4, 17])
def test_split_layouts(part, num):
    k = oracle.prng_key(20241017)
    out = oracle.split(k, num, part)
    assert np.array_equal(out, dfx.random.split(k, num, partitionable=part))
    if part:   # foldlike: child i = block(key, (0, i))
        for i in range(num):
            assert tuple(out[i]) == oracle.threefry2x32(int(k[0]), int(k[1]), 0, i)
    else:      # original: counts iota(2n) halved; outputs concatenated then reshaped (n, 2)
        flat = np.zeros(2 * num, np.uint32)
        for i in range(num):
            flat[i], flat[num + i] = oracle.threefry2x32(int(k[0]), int(k[1]), i, num + i)
        assert np.array_equal(out, flat.reshape(num, 2))
    assert len({tuple(r) for r in out}) == num


def test_erfinv_against_scipy():
    xs = np.linspace(-0.999, 0.999, 20001)
    e64 = np.array([oracle.erfinv(x) for x in xs])
    ref = sp.erfinv(xs)
    assert np.max(np.abs(e64 - ref) / np.maximum(np.abs(ref), 1e-300)) < 5e-15
    xs32 = xs.astype(np.float32)
    e32 = np.array([oracle.erfinv(x, np.float32) for x in xs32], np.float64)
    ref32 = sp.erfinv(xs32.astype(np.float64))
    assert np.nanmax(np.abs(e32 - ref32) / np.maximum(np.abs(ref32), 1e-30)) < 1e-5   # Giles SP: ~6e-6 relative
    # tails: inverse property erf(erfinv(x)) == x to rounding
    for x in (1 - 1e-9, 1 - 1e-12, -(1 - 1e-15)):
        assert abs(sp.erf(oracle.erfinv(x)) - x) < 3e-16
    assert oracle.erfinv(1.0) == np.inf and oracle.erfinv(-1.0) == -np.inf


def test_erfinv_tails_against_mpmath():
    mp = pytest.importorskip("mpmath")
    mp.mp.dps = 40
    # the XLA form w = -log1p(-x*x) loses accuracy as e^w * eps towards |x| -> 1; the polynomial itself is ~1e-15
    for x, tol in ((0.9989568, 1e-14), (0.99999, 1e-13), (1 - 2.0 ** -30, 2e-9)):
        r = mp.erfinv(mp.mpf(x))
        assert abs((mp.mpf(oracle.erfinv(x)) - r) / r) < tol


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("part", [True, False])
def test_normal_is_standard_normal(dtype, part):
    from scipy import stats
    keys = oracle.split(oracle.prng_key(3), 20000, part)
    z = np.array([oracle.normal(k, dtype, part) for k in keys], np.float64)
    assert np.all(np.isfinite(z))
    assert stats.kstest(z, "norm").pvalue > 0.01
    assert abs(z.mean()) < 0.03 and abs(z.var() - 1) < 0.04


def test_normal_uniform_map_bits():
    """mantissa fill: bits>>9 | 0x3F800000 -> [1,2) - 1 -> u = max(lo, 2f + lo) in [lo, 1)."""
    k = oracle.prng_key(5)
    for part in (True, False):
        a, b = oracle.threefry2x32(int(k[0]), int(k[1]), 0, 0)
        bits = (a ^ b) if part else a
        f = np.array([(bits >> 9) | 0x3F800000], np.uint32).view(np.float32)[0] - np.float32(1)
        lo = np.nextafter(np.float32(-1), np.float32(0))
        u = max(lo, np.float32(f * (np.float32(1) - lo) + lo))
        want = np.float32(np.sqrt(2)) * np.float32(sp.erfinv(np.float64(u)))
        got = oracle.normal(k, np.float32, part)
        assert abs(float(got) - float(want)) <= 2e-5 * max(1.0, abs(float(want)))
