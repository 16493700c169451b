"""Parity against the reference ITSELF, whenever it can be had:

* a live Diffrax importable right now (``baseline.probe()``): every golden case is run through
  ``diffrax.diffeqsolve`` under ``jax.vmap`` and the oracle must agree at the north-star tolerances - Brownian / PRNG words
  bit-exact on the integer side, accepted-step counts equal or +-1, saved states within 1e-10 (fp64) / 1e-4 (fp32);
* or ``tests/golden/diffrax_golden.npz`` written earlier by ``baseline/gen_golden.py``.

Neither exists in the authoring container (no jax wheel, no network): the tests then SKIP with the probe's reason, and
the repository's parity status stays "unpinned against a live Diffrax".  The probe itself is always tested.
"""
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))

import baseline  # noqa: E402
import make_golden  # noqa: E402
import oracle  # noqa: E402

GOLD_PATH = os.path.join(HERE, "golden", "diffrax_golden.npz")
OK, WHY = baseline.probe()


def test_probe_reports_a_reason():
    ok, why = baseline.probe()
    assert isinstance(ok, bool) and isinstance(why, str) and why
    v = baseline.versions()
    assert v["available"] == ok
    if not ok:
        with pytest.raises(RuntimeError):
            baseline.solve(make_golden.CASES["c2_lorenz_dopri5_t1"])


def _reference(name):
    if OK:
        return baseline.solve(make_golden.CASES[name])
    if os.path.exists(GOLD_PATH):
        g = np.load(GOLD_PATH)
        if f"{name}/ys" in g:
            return {k: g[f"{name}/{k}"] for k in ("ys", "ts", "stats", "result")}
    pytest.skip(f"no live Diffrax ({WHY}) and no tests/golden/diffrax_golden.npz")


def _rel(a, b):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    m = np.isfinite(b)
    assert np.array_equal(np.isfinite(a), m)
    return float(np.max(np.abs(a[m] - b[m]) / (np.abs(b[m]) + 1e-3 * np.abs(b[m]).max()))) if m.any() else 0.0


@pytest.mark.parametrize("name", [n for n in make_golden.CASES if "dense" not in n])
def test_oracle_against_live_diffrax(name):
    ref = _reference(name)
    kw = dict(make_golden.CASES[name])
    field = kw.pop("field")
    y0, t0, t1, dt0 = kw.pop("y0"), kw.pop("t0"), kw.pop("t1"), kw.pop("dt0")
    o = oracle.solve(field, y0, t0, t1, dt0, **kw)
    f32 = np.dtype(kw.get("dtype", np.float64)) == np.float32
    assert np.abs(o["stats"][:, 1] - ref["stats"][:, 1]).max() <= 1          # accepted steps equal or +-1
    same = np.all(o["stats"] == ref["stats"], axis=1)
    ys = ref["ys"].reshape(o["ys"].shape)
    assert _rel(o["ys"][same], ys[same]) < (1e-4 if f32 else 1e-10)
    assert np.array_equal(o["result"][same], ref["result"][same])


def test_prng_words_against_jax():
    if not (OK or os.path.exists(GOLD_PATH)):
        pytest.skip(f"no live Diffrax ({WHY}) and no tests/golden/diffrax_golden.npz")
    if OK and not os.path.exists(GOLD_PATH):
        pytest.skip("run baseline/gen_golden.py to write the jax.random words")
    g = np.load(GOLD_PATH)
    keys = g["prng/keys"]
    for part in (1, 0):
        assert np.array_equal(np.stack([oracle.split(k, 3, bool(part)) for k in keys[:16]]), g[f"prng/split3_part{part}"])
        for dt_, tag, ulp in ((np.float64, "f64", 4), (np.float32, "f32", 4)):
            z = np.array([oracle.normal(k, dt_, bool(part)) for k in keys], dt_)
            ref = g[f"prng/normal_{tag}_part{part}"]
            # integer side exact <=> the uniforms coincide <=> the normals agree to the float side's last ulps
            assert np.max(np.abs(z.astype(np.float64) - ref.astype(np.float64)) / np.spacing(np.abs(ref)).astype(np.float64)) <= ulp
            z3 = np.stack([oracle.normal(k, dt_, bool(part), (3,)) for k in keys])
            ref3 = g[f"prng/normal3_{tag}_part{part}"]
            assert np.max(np.abs(z3.astype(np.float64) - ref3.astype(np.float64)) / np.spacing(np.abs(ref3)).astype(np.float64)) <= ulp
