"""The oracle is regression-pinned to the committed golden vectors (tests/golden/make_golden.py)."""
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import make_golden  # noqa: E402
import oracle  # noqa: E402

GOLD = np.load(os.path.join(HERE, "golden", "ensemble_golden.npz"))


@pytest.mark.parametrize("name", list(make_golden.CASES) + list(make_golden.CASES_EXTRA))
def test_oracle_reproduces_golden(name):
    kw = dict(make_golden.CASES[name] if name in make_golden.CASES else make_golden.CASES_EXTRA[name])
    field = kw.pop("field")
    y0, t0, t1, dt0 = kw.pop("y0"), kw.pop("t0"), kw.pop("t1"), kw.pop("dt0")
    r = oracle.solve(field, y0, t0, t1, dt0, **kw)
    assert np.array_equal(r["stats"], GOLD[f"{name}/stats"])
    assert np.array_equal(r["result"], GOLD[f"{name}/result"])
    # same binary on the same libm is bit-identical; allow last-ulp libm differences across hosts
    assert np.allclose(r["ys"], GOLD[f"{name}/ys"], rtol=1e-12 if r["ys"].dtype == np.float64 else 1e-5, atol=0, equal_nan=True)
    assert np.array_equal(np.isinf(r["ts"]), np.isinf(GOLD[f"{name}/ts"]))


def test_arenstorf_orbit_is_periodic():
    """Known answer for the CR3BP field + Dopri8: the Arenstorf orbit closes after one period."""
    y0 = np.array([[0.994, 0.0, 0.0, -2.00158510637908252]])
    r = oracle.solve("cr3bp", y0, 0.0, 17.0652165601579625, None, solver="dopri8", params=[0.012277471],
                     rtol=1e-12, atol=1e-12)
    assert np.allclose(r["ys"][0, 0], y0[0], atol=1e-7)
