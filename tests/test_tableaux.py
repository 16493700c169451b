"""Butcher tableaux: the generated coefficient tables against the checks the reference applies
to its own tableaux (runge_kutta.py:126-130, 143-163) and against the Butcher
order conditions, so a wrong coefficient cannot hide in both the oracle and the kernels."""
import itertools
import json
import os
import re

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
TABS = json.load(open(os.path.join(HERE, "golden", "tableaux.json")))
ORDERS = {"tsit5": 5, "dopri5": 5, "dopri8": 8, "heun": 2, "bosh3": 3, "midpoint": 2, "ralston": 2}
EMBEDDED = {"tsit5": 4, "dopri5": 4, "dopri8": 7, "heun": 1, "bosh3": 2, "midpoint": 1, "ralston": 1}
FSAL = {"tsit5": True, "dopri5": True, "dopri8": True, "heun": False, "bosh3": True, "midpoint": False, "ralston": False}


def fx(h):
    return float.fromhex(h)


def dense(name):
    t = TABS[name]
    b = np.array([fx(x) for x in t["b_sol"]])
    s = len(b)
    A = np.zeros((s, s))
    for i, row in enumerate(t["a_lower"]):
        A[i + 1, : i + 1] = [fx(x) for x in row]
    c = np.array([0.0] + [fx(x) for x in t["c"]])
    be = np.array([fx(x) for x in t["b_error"]])
    return A, b, be, c


@pytest.mark.parametrize("name", list(ORDERS))
def test_reference_post_init_checks(name):
    A, b, be, c = dense(name)
    assert np.allclose(A.sum(1), c)                      # runge_kutta.py:126-128
    assert np.allclose(b.sum(), 1.0)                     # 129
    assert np.allclose(be.sum(), 0.0)                    # 130
    fsal = bool((b[:-1] == A[-1, :-1]).all() and b[-1] == 0.0)   # 143-157
    assert fsal == FSAL[name]


def _trees(order):
    """Rooted trees up to `order` as nested tuples, with density gamma and elementary weight."""
    by_order = {1: [()]}
    for n in range(2, order + 1):
        out = set()
        for parts in _partitions(n - 1):
            for combo in itertools.product(*[by_order[p] for p in parts]):
                out.add(tuple(sorted(combo, key=repr)))
        by_order[n] = sorted(out, key=repr)
    return by_order


def _partitions(n, maxp=None):
    maxp = maxp or n
    if n == 0:
        yield ()
        return
    for k in range(min(n, maxp), 0, -1):
        for rest in _partitions(n - k, k):
            yield (k,) + rest


def _order_of(t):
    return 1 + sum(_order_of(c) for c in t)


def _gamma(t):
    g = _order_of(t)
    for c in t:
        g *= _gamma(c)
    return g


def _phi(t, A, s):
    """Vector of elementary weights per stage: Phi_i(t) = prod_children (A @ Phi(child))_i."""
    v = np.ones(s)
    for c in t:
        v = v * (A @ _phi(c, A, s))
    return v


@pytest.mark.parametrize("name", list(ORDERS))
def test_order_conditions(name):
    A, b, be, c = dense(name)
    s = len(b)
    p = ORDERS[name]
    trees = _trees(min(p, 6))  # all rooted trees up to order 6 (37 trees); higher orders via quadrature below
    for n in range(1, min(p, 6) + 1):
        for t in trees[n]:
            lhs = b @ _phi(t, A, s)
            assert abs(lhs - 1.0 / _gamma(t)) < 1e-9, (name, n, t, lhs)
    # embedded solution b - b_error has order EMBEDDED[name]
    bh = b - be
    q = EMBEDDED[name]
    for n in range(1, min(q, 6) + 1):
        for t in trees[n]:
            assert abs(bh @ _phi(t, A, s) - 1.0 / _gamma(t)) < 1e-9, (name, "embedded", n, t)
    # bushy-tree (quadrature) conditions up to the full order: b . c^(k-1) = 1/k
    for k in range(1, p + 1):
        assert abs(b @ c ** (k - 1) - 1.0 / k) < 5e-10, (name, k)


def test_appendix_c_facts():
    assert fx(TABS["tsit5"]["b_error"][-1]) == -1 / 66           # tsit5.py:87
    assert fx(TABS["dopri5"]["b_error"][-1]) == -1.0 / 60.0      # dopri5.py:29
    assert fx(TABS["dopri8"]["b_error"][-2]) == 1 / 4 and fx(TABS["dopri8"]["b_error"][-1]) == 0.0
    assert fx(TABS["dopri5"]["b_error"][0]) == 35 / 384 - 1951 / 21600  # differenced in double, not in fp32
    sh = TABS["shark"]
    assert [fx(x) for x in sh["b_sol"]] == [0.4, 0.6] and fx(sh["a"][0][0]) == 5 / 6


def test_generated_headers_match_golden():
    """oracle/oracle_tableaux.h and csrc/tableaux.cuh are regenerable bit-for-bit from the golden JSON."""
    import subprocess, sys, tempfile, shutil
    hdrs = [os.path.join(ROOT, "oracle", "oracle_tableaux.h"), os.path.join(ROOT, "diffrax_b200", "csrc", "tableaux.cuh")]
    before = [open(h).read() for h in hdrs]
    for h, txt in zip(hdrs, before):
        for hexv in re.findall(r"-?0x1\.[0-9a-f]+p[+-]\d+", txt)[:50]:
            float.fromhex(hexv)
    # every golden Dopri5 a-coefficient appears verbatim in both headers
    for row in TABS["dopri5"]["a_lower"]:
        for hv in row:
            lit = "0.0" if fx(hv) == 0.0 else float(fx(hv)).hex()
            assert all(lit in txt for txt in before)


def test_golden_matches_reference_if_present():
    """When the reference tree is mounted (authoring container only) re-extract and compare."""
    ref = "/root/reference/diffrax/_solver/dopri5.py"
    if not os.path.exists(ref):
        pytest.skip("reference tree not mounted (GPU box)")
    import subprocess, sys, tempfile
    src = open(os.path.join(ROOT, "tools", "extract_tableaux.py")).read()
    with tempfile.TemporaryDirectory() as td:
        os.makedirs(os.path.join(td, "tools")); os.makedirs(os.path.join(td, "tests", "golden"))
        open(os.path.join(td, "tools", "extract_tableaux.py"), "w").write(src)
        subprocess.run([sys.executable, os.path.join(td, "tools", "extract_tableaux.py")], check=True, capture_output=True)
        assert json.load(open(os.path.join(td, "tests", "golden", "tableaux.json"))) == TABS
