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


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_sequenced_log1p_within_one_ulp(dtype):
    """The explicitly sequenced log1p (oracle.c / csrc/prng.cuh; domain [-1, 0], all erf_inv needs) against mpmath: <= 1 ulp."""
    mp = pytest.importorskip("mpmath")
    mp.mp.prec = 200
    rng = np.random.default_rng(0)
    tiny = -7 if dtype == np.float32 else -15
    xs = np.concatenate([-rng.uniform(0, 1, 3000) ** 2, -10.0 ** rng.uniform(-18, 0, 1500), -rng.uniform(0.28, 0.31, 1000),
                         -rng.uniform(0.49, 0.51, 500), -1 + 10.0 ** rng.uniform(tiny, 0, 1000),
                         [-0.2928932188134524, -0.2928932188134525, -0.5, -0.75, 0.0]]).astype(dtype)
    xs = xs[(xs > -1) & (xs <= 0)]
    got = oracle.log1p(xs, dtype).astype(np.float64)
    ref = np.array([float(mp.log1p(mp.mpf(float(v)))) for v in xs])
    ulp = np.spacing(np.abs(ref.astype(dtype))).astype(np.float64)
    assert np.max(np.abs(got - ref) / np.where(ulp > 0, ulp, 1.0)) <= 1.0
    assert oracle.log1p(np.array([0.0], dtype), dtype)[0] == 0.0
    assert oracle.log1p(np.array([-1.0], dtype), dtype)[0] == -np.inf
    assert np.isnan(oracle.log1p(np.array([-1.5], dtype), dtype)[0]) and np.isnan(oracle.log1p(np.array([0.5], dtype), dtype)[0])


def _jax_random_bits(key, bit_width, m, part):
    """jax/_src/prng.py threefry_random_bits for shape (m,), written whole-array the way JAX does it (iota, pad, halve,
    concatenate) - an independent statement of the element <-> counter layout the oracle indexes per element."""
    k0, k1 = int(key[0]), int(key[1])
    tf = dfx.random.threefry2x32
    if part:  # _threefry_random_bits_partitionable: counters (hi, lo) = (0, flat index)
        b1, b2 = tf(k0, k1, np.zeros(m, np.uint32), np.arange(m, dtype=np.uint32))
        return (b1.astype(np.uint64) << np.uint64(32)) | b2.astype(np.uint64) if bit_width == 64 else b1 ^ b2
    max_count = int(np.ceil(bit_width * m / 32))            # _threefry_random_bits_original
    count = np.arange(max_count, dtype=np.uint32)
    odd = count.size % 2                                    # threefry_2x32: pad to even, split in halves, concatenate
    if odd:
        count = np.concatenate([count, np.zeros(1, np.uint32)])
    x0, x1 = np.split(count, 2)
    y0, y1 = tf(k0, k1, x0, x1)
    bits = np.concatenate([y0, y1])
    bits = bits[:-1] if odd else bits
    if bit_width == 64:
        hi, lo = np.split(bits, 2)
        return (hi.astype(np.uint64) << np.uint64(32)) | lo.astype(np.uint64)
    return bits


@pytest.mark.parametrize("part", [True, False])
@pytest.mark.parametrize("m", [1, 2, 3, 4, 5, 8])
def test_vector_normal_layout(part, m):
    """jr.normal(key, (m,)): element w of the oracle's per-element indexing == the whole-array JAX layout; m == 1 is the
    scalar draw.  Checked on the integer side through the (monotone, injective on the mantissa field) uniform map."""
    for seed in (1, 99):
        k = oracle.prng_key(seed)
        for dtype, width in ((np.float32, 32), (np.float64, 64)):
            bits = _jax_random_bits(k, width, m, part)
            got = oracle.normal(k, dtype, part, (m,))
            # rebuild the uniforms from the whole-array bits and push them through the oracle's own float side
            for w in range(m):
                # a scalar key whose block (0,0)/(0,1) we cannot choose, so compare via the uniform's mantissa: invert normal -> u
                if width == 32:
                    f = np.array([(int(bits[w]) >> 9) | 0x3F800000], np.uint32).view(np.float32)[0] - np.float32(1)
                    lo = np.nextafter(np.float32(-1), np.float32(0))
                    u = max(lo, np.float32(f * np.float32(2) + lo))
                    want = np.float32(np.sqrt(2)) * oracle.erfinv(u, np.float32)
                else:
                    f = np.array([(int(bits[w]) >> 12) | 0x3FF0000000000000], np.uint64).view(np.float64)[0] - 1.0
                    lo = np.nextafter(-1.0, 0.0)
                    u = max(lo, f * 2.0 + lo)
                    want = np.sqrt(2.0) * oracle.erfinv(u)
                assert got[w] == want or abs(got[w] - want) <= 2 * np.spacing(abs(want))  # (fma vs mul+add in the uniform map)
        if m == 1:
            assert oracle.normal(k, np.float64, part, (1,))[0] == oracle.normal(k, np.float64, part)
            assert oracle.normal(k, np.float32, part, (1,))[0] == oracle.normal(k, np.float32, part)
