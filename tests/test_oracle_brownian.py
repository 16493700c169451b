"""Pins the oracle's VirtualBrownianTree with the reference's own (statistical) Brownian tests:
test_brownian.py:117-180 (increments ~ N(0, t1-t0), H ~ N(0, (t1-t0)/12), W _|_ H), 460-490
(conditional statistics of the bridge), 679-695 (reverse-time antisymmetry) and the SDE strong
order checks of test_sde1.py:17-94 / test_integrate.py:193-322."""
import numpy as np
import pytest
from scipy import stats

import oracle

N = 60000  # the reference uses 600 000 vmapped keys; 60 000 keeps the CPU suite fast


@pytest.fixture(scope="module")
def keys():
    return oracle.split(oracle.prng_key(1234), N)


@pytest.mark.parametrize("levy", ["bi", "stla"])
@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("interval", [(0.0, 3.0), (0.3, 2.2), (1.0, 1.0 + 2 ** -9)])
def test_increment_statistics(keys, levy, dtype, interval):
    ta, tb = interval
    W, H = oracle.vbt_evaluate(keys, ta, tb, bm_t0=0.0, bm_t1=3.0, tol=2.0 ** -7, levy_area=levy, dtype=dtype)
    dt = tb - ta
    assert stats.kstest(W.astype(np.float64) / np.sqrt(dt), "norm").pvalue > 0.01
    if levy == "stla":
        assert stats.kstest(H.astype(np.float64) / np.sqrt(dt / 12), "norm").pvalue > 0.01
        assert abs(np.mean(W.astype(np.float64) * H.astype(np.float64))) < 0.02 * dt   # test_brownian.py:176


@pytest.mark.parametrize("levy", ["bi", "stla"])
def test_chen_additivity_and_independence(keys, levy):
    kw = dict(bm_t0=0.0, bm_t1=1.0, tol=2.0 ** -8, levy_area=levy)
    W1, H1 = oracle.vbt_evaluate(keys, 0.0, 0.4, **kw)
    W2, H2 = oracle.vbt_evaluate(keys, 0.4, 1.0, **kw)
    W3, H3 = oracle.vbt_evaluate(keys, 0.0, 1.0, **kw)
    assert np.max(np.abs(W1 + W2 - W3)) < 1e-14
    assert abs(np.corrcoef(W1, W2)[0, 1]) < 0.02
    if levy == "stla":
        # Chen's relation for space-time Levy area over [s,u] = [s,t] U [t,u]
        s, t, u = 0.0, 0.4, 1.0
        H_chen = ((t - s) * H1 + (u - t) * H2 + 0.5 * ((u - t) * W1 - (t - s) * W2)) / (u - s)
        assert np.max(np.abs(H_chen - H3)) < 1e-13


def test_reverse_time_antisymmetry(keys):
    # test_brownian.py:679-695: W(t1, t0) == -W(t0, t1)
    W, _ = oracle.vbt_evaluate(keys[:1000], 0.2, 0.9, tol=2.0 ** -8)
    Wr, _ = oracle.vbt_evaluate(keys[:1000], 0.9, 0.2, tol=2.0 ** -8)
    assert np.array_equal(W, -Wr)


def test_conditional_bridge_statistics(keys):
    # test_brownian.py:460-490: W_r | W_s, W_u  ~ N(W_s + (r-s)/(u-s) (W_u - W_s), (r-s)(u-r)/(u-s))
    s, r, u = 0.25, 0.40625, 0.5
    Ws, _ = oracle.vbt_evaluate(keys, 0.0, s, tol=2.0 ** -6)
    Wr, _ = oracle.vbt_evaluate(keys, 0.0, r, tol=2.0 ** -6)
    Wu, _ = oracle.vbt_evaluate(keys, 0.0, u, tol=2.0 ** -6)
    resid = Wr - (Ws + (r - s) / (u - s) * (Wu - Ws))
    var = (r - s) * (u - r) / (u - s)
    assert stats.kstest(resid / np.sqrt(var), "norm").pvalue > 0.01
    assert abs(np.corrcoef(resid, Wu - Ws)[0, 1]) < 0.02


def test_descent_key_selection_quirk():
    """Going RIGHT (r > t_mid) continues with key_st, LEFT with key_tu (tree.py:431-432): re-derive W(1/4) by hand."""
    k = oracle.prng_key(77)
    leaf = oracle.split(k, 1)[0]
    state, kw = oracle.split(leaf, 2)
    W1 = oracle.normal(kw)
    # level 0: interval [0,1], midpoint 1/2; r = 0.25 goes LEFT -> key_tu
    key_st, mid, key_tu = oracle.split(state, 3)
    w_half = 0.5 * W1 + (np.sqrt(1.0) / 2) * oracle.normal(mid)
    # level 1: interval [0,1/2], midpoint 1/4; r = 0.25 is not > 0.25 -> LEFT again
    key_st2, mid2, key_tu2 = oracle.split(key_tu, 3)
    w_quarter = 0.5 * w_half + (np.sqrt(0.5) / 2) * oracle.normal(mid2)
    # tol = 1/4 -> two levels; final interval [0, 1/4], sr = 1/4, ru = 0 -> bridge term vanishes
    W, _ = oracle.vbt_evaluate(k[None], 0.0, 0.25, tol=0.25)
    assert abs(W[0] - w_quarter) < 1e-15


@pytest.mark.parametrize("solver,levy,order", [("euler", "bi", 1.0), ("heun", "bi", 1.0), ("shark", "stla", 1.5)])
def test_sde_strong_order_additive_noise(solver, levy, order):
    """OU (additive noise): Euler/Heun strong order 1 for additive noise, ShARK 1.5 (shark.py:57-64).
    Reference = same Brownian paths at a much finer step (test/helpers.py:136-295 recipe)."""
    nk = 200
    keys = oracle.split(oracle.prng_key(99), nk)
    kw = dict(params=[1.0, 0.0, 0.5], controller="constant", levy_area=levy, keys=keys, bm_tol=2.0 ** -14)
    fine = oracle.solve("ou", np.ones((nk, 1)), 0.0, 1.0, 2.0 ** -11, solver="shark", max_steps=1 << 14,
                        **{**kw, "levy_area": "stla"})["ys"][:, 0, 0]
    if levy == "bi":  # a BrownianIncrement tree is a different path than the STLA tree: build its own fine reference
        fine = oracle.solve("ou", np.ones((nk, 1)), 0.0, 1.0, 2.0 ** -11, solver="heun", max_steps=1 << 14, **kw)["ys"][:, 0, 0]
    errs, dts = [], []
    for k in range(2, 7):
        dt = 2.0 ** -k
        y = oracle.solve("ou", np.ones((nk, 1)), 0.0, 1.0, dt, solver=solver, **kw)["ys"][:, 0, 0]
        errs.append(np.sqrt(np.mean((y - fine) ** 2))); dts.append(dt)
    slope = np.polyfit(np.log(dts), np.log(errs), 1)[0]
    assert slope > order - 0.35, (solver, slope, errs)


def test_ou_moments():
    n = 40000
    keys = oracle.split(oracle.prng_key(5), n)
    for solver, levy in (("heun", "bi"), ("shark", "stla")):
        r = oracle.solve("ou", np.ones((n, 1)), 0.0, 1.0, 2.0 ** -6, solver=solver, params=[1.0, 0.0, 0.5],
                         controller="constant", levy_area=levy, keys=keys, bm_tol=2.0 ** -8)
        y = r["ys"][:, 0, 0]
        assert abs(y.mean() - np.exp(-1)) < 4 * np.sqrt(0.108 / n) + 2e-3
        assert abs(y.var() - 0.125 * (1 - np.exp(-2))) < 3e-3
        assert np.all(r["stats"][:, 0] == 64)


def test_vector_brownian_motion_leaves():
    """VirtualBrownianTree(shape=(m,)) is ONE leaf (a tuple of ints becomes a single ShapeDtypeStruct, tree.py:291-295) whose
    key is split_by_tree(key, shape) = jr.split(key, 1)[0] (tree.py:301, _misc.py:128-133); every node draws
    jr.normal(key, (m,)).  With partitionable threefry element 0 of an (m,) draw uses counter (0, 0) like the scalar draw, so
    component 0 of a diagonal-noise solve is the scalar solve; with the legacy layout the counters depend on m and it is
    not.  Components are independent."""
    import diffrax_b200 as dfx
    n = 64
    keys = dfx.random.split(dfx.random.key(21), n)
    common = dict(params=[1.0, 0.0, 0.5], controller="constant", keys=keys, bm_tol=2.0 ** -8, solver="heun", levy_area="bi")
    for part in (True, False):
        vec = oracle.solve("ou", np.ones((n, 3)), 0.0, 1.0, 2.0 ** -5, bm_dim=3, partitionable=part, **common)
        sca = oracle.solve("ou", np.ones((n, 1)), 0.0, 1.0, 2.0 ** -5, partitionable=part, **common)
        same0 = np.array_equal(vec["ys"][:, -1, 0], sca["ys"][:, -1, 0])
        assert same0 == part
        c = np.corrcoef(vec["ys"][:, -1, :].T)
        assert abs(c[0, 1]) < 0.4 and abs(c[0, 2]) < 0.4 and abs(c[1, 2]) < 0.4
        # exact OU law per component: mean e^-1, variance sigma^2 (1 - e^-2) / 2
        assert abs(vec["ys"][:, -1, :].mean() - np.exp(-1.0)) < 0.1


@pytest.mark.parametrize("levy", ["bi", "stla"])
def test_vector_tree_shares_the_key_path(levy):
    """shape (m,) vs shape (): same leaf key, same descent keys; only the element of each normal draw differs.  So the
    (m,) increment's statistics per component match the scalar law, W of a 3-vector has independent N(0, t-s) components,
    and (partitionable) component 0 equals the scalar tree bit for bit."""
    import diffrax_b200 as dfx
    keys = dfx.random.split(dfx.random.key(4), 4000)
    W3, H3 = oracle.vbt_evaluate(keys, 0.2, 0.9, tol=2.0 ** -8, levy_area=levy, shape=(3,))
    W1, H1 = oracle.vbt_evaluate(keys, 0.2, 0.9, tol=2.0 ** -8, levy_area=levy)
    assert W3.shape == (4000, 3)
    assert np.array_equal(W3[:, 0], W1) and np.array_equal(H3[:, 0], H1)
    assert np.all(np.abs(W3.var(0) - 0.7) < 0.06) and np.all(np.abs(W3.mean(0)) < 0.05)
    c = np.corrcoef(W3.T)
    assert abs(c[0, 1]) < 0.06 and abs(c[0, 2]) < 0.06 and abs(c[1, 2]) < 0.06
    if levy == "stla":  # H ~ N(0, (t-s)/12), independent of W
        assert np.all(np.abs(H3.var(0) - 0.7 / 12) < 0.01)
        assert abs(np.corrcoef(W3[:, 1], H3[:, 1])[0, 1]) < 0.06


@pytest.mark.parametrize("solver,levy", [("euler", "bi"), ("heun", "bi"), ("shark", "stla")])
def test_matrix_valued_diffusion_covariance_law(solver, levy):
    """General ControlTerm: a constant [d, m] diffusion matrix G with an m-dimensional Brownian motion, prod = tensordot(G, dW)
    (_term.py:267-268, 417-427).  dy = theta (mu - y) dt + G dW has the exact law  mean mu + (y0 - mu) e^{-theta t},
    covariance G G^T (1 - e^{-2 theta t}) / (2 theta): the oracle's ensemble reproduces it (correlated components)."""
    import diffrax_b200 as dfx
    n = 6000
    keys = dfx.random.split(dfx.random.key(5), n)
    G = np.array([[0.5, 0.2], [0.0, 0.3], [0.1, -0.4]])
    r = oracle.solve(oracle.FIELDS["ou_matrix2"], np.ones((n, 3)), 0.0, 1.0, 2.0 ** -6, solver=solver, params=[1.0, 0.0] + list(G.ravel()),
                     controller="constant", levy_area=levy, keys=keys, bm_tol=2.0 ** -8, bm_dim=2)
    y = r["ys"][:, -1, :]
    want = G @ G.T * (1 - np.exp(-2.0)) / 2
    tol = 0.012 if solver != "euler" else 0.02
    assert np.abs(np.cov(y.T) - want).max() < tol
    assert np.abs(y.mean(0) - np.exp(-1.0)).max() < 0.02


def test_state_dependent_diffusion_gbm_converges_to_the_exact_solutions():
    """ControlTerm with a state-dependent diffusion g(t, y) = sigma y (geometric Brownian motion), scalar Brownian motion from
    the VirtualBrownianTree.  Path-wise against the closed forms driven by THE SAME W(1) (oracle.vbt_evaluate):
    Heun -> Stratonovich  y0 exp(mu t + sigma W)   (heun.py:24-33),   Euler -> Ito  y0 exp((mu - sigma^2 / 2) t + sigma W)."""
    import diffrax_b200 as dfx
    n = 400
    keys = dfx.random.split(dfx.random.key(8), n)
    mu, sigma = 0.3, 0.4
    W, _ = oracle.vbt_evaluate(keys, 0.0, 1.0, tol=2.0 ** -12)
    errs = {}
    for solver, exact in (("heun", np.exp(mu + sigma * W)), ("euler", np.exp(mu - 0.5 * sigma ** 2 + sigma * W))):
        errs[solver] = []
        for k in (6, 8, 10):
            r = oracle.solve("gbm", np.ones((n, 1)), 0.0, 1.0, 2.0 ** -k, solver=solver, params=[mu, sigma], controller="constant",
                             levy_area="bi", keys=keys, bm_tol=2.0 ** -12, max_steps=1 << 12)
            errs[solver].append(np.sqrt(np.mean((r["ys"][:, -1, 0] - exact) ** 2)))
    assert errs["heun"][-1] < 2e-4 and errs["euler"][-1] < 1.5e-2
    # strong order: commutative noise - Heun 1.0, Euler-Maruyama 0.5; error ratio over a factor 16 in dt
    assert errs["heun"][0] / errs["heun"][-1] > 8.0
    assert 2.5 < errs["euler"][0] / errs["euler"][-1] < 7.0
