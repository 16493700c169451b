"""Pins the oracle's ODE path against the anchors the reference's own tests use:
analytic solutions (test_integrate.py:48-141, test_saveat_solution.py:24-195), convergence order
(test_integrate.py:144-190), scipy DOP853 on DETEST-style problems (test_detest.py:390-469),
SaveAt semantics (test_saveat_solution.py:116-180, 323-428), reverse time (test_integrate.py:325-413)."""
import math

import numpy as np
import pytest
from scipy.integrate import solve_ivp

import oracle

ADAPTIVE = ["tsit5", "dopri5", "dopri8", "heun", "bosh3", "midpoint", "ralston"]


@pytest.mark.parametrize("solver", ADAPTIVE)
@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_basic_decay(solver, dtype):
    # test_integrate.py:48-141: dy=-y, y(1) = y0/e within 1e-2
    y0 = np.array([[2.1], [0.3]])
    r = oracle.solve("decay", y0, 0.0, 1.0, 0.01, solver=solver, params=[1.0], dtype=dtype, rtol=1e-4, atol=1e-6)
    assert np.all(r["result"] == 0)
    assert np.allclose(r["ys"][:, 0, 0], y0[:, 0] / math.e, rtol=1e-2, atol=1e-2)
    assert r["ys"].dtype == np.dtype(dtype)


def test_saveat_ts_matches_analytic():
    # test_saveat_solution.py:24-41, 97-114: dy=-0.5y, Dopri5, PID(1e-8,1e-8), dt0=None, ts=[0.5,0.8]
    r = oracle.solve("decay", np.array([[2.1]]), 0.0, 1.0, None, solver="dopri5", params=[0.5], rtol=1e-8, atol=1e-8,
                     save_t1=False, save_ts=[0.5, 0.8])
    assert np.array_equal(r["ts"][0], [0.5, 0.8])
    assert np.allclose(r["ys"][0, :, 0], 2.1 * np.exp(-0.5 * np.array([0.5, 0.8])), rtol=1e-5, atol=1e-8)
    assert r["stats"][0, 0] > 0


def test_saveat_steps_counts_and_padding():
    # test_saveat_solution.py:116-180
    kw = dict(solver="dopri5", params=[0.5], rtol=1e-8, atol=1e-8)
    r = oracle.solve("decay", np.array([[2.1]]), 0.0, 1.0, None, save_t1=False, save_steps=1, **kw)
    assert r["ts"].shape == (1, 4096) and r["ys"].shape == (1, 4096, 1)
    n = r["stats"][0, 1]
    assert np.all(np.isfinite(r["ts"][0, :n])) and np.all(r["ts"][0, n:] == np.inf) and np.all(r["ys"][0, n:] == np.inf)
    assert r["ts"][0, n - 1] == 1.0
    assert np.allclose(r["ys"][0, :n, 0], 2.1 * np.exp(-0.5 * r["ts"][0, :n]), rtol=1e-5, atol=1e-8)
    r2 = oracle.solve("decay", np.array([[2.1]]), 0.0, 1.0, None, save_t1=False, save_steps=2, **kw)
    assert r2["ts"].shape[1] == 4096 // 2
    n2 = np.isfinite(r2["ts"][0]).sum()
    assert n2 == n // 2 and np.array_equal(r2["ts"][0, :n2], r["ts"][0, 1:n:2])
    r3 = oracle.solve("decay", np.array([[2.1]]), 0.0, 1.0, None, save_t1=True, save_steps=2, **kw)
    assert r3["ts"].shape[1] == 4096 // 2        # max_steps % 2 == 0 -> no extra slot (1288-1293)
    r4 = oracle.solve("decay", np.array([[2.1]]), 0.0, 1.0, None, save_t1=True, save_steps=3, **kw)
    assert r4["ts"].shape[1] == 4096 // 3 + 1
    n4 = np.isfinite(r4["ts"][0]).sum()
    assert r4["ts"][0, n4 - 1] == 1.0 and n4 == n // 3 + (1 if n % 3 else 0)
    r5 = oracle.solve("decay", np.array([[2.1]]), 0.0, 1.0, None, save_t0=True, save_t1=True, save_ts=[0.25, 0.75], **kw)
    assert np.array_equal(r5["ts"][0], [0.0, 0.25, 0.75, 1.0])
    assert r5["ys"][0, 0, 0] == 2.1


def test_t0_equals_t1():
    # test_saveat_solution.py:323-428: every requested output returns y0; per-lane under "vmap"
    y0 = np.array([[2.1], [0.7]])
    r = oracle.solve("decay", y0, 0.0, 0.0, 0.1, solver="tsit5", params=[1.0], rtol=1e-6, atol=1e-6,
                     save_t0=True, save_t1=True, save_ts=[0.0, 0.0], t0_per_traj=[0.0, 0.0], t1_per_traj=[0.0, 1.0])
    assert np.array_equal(r["ys"][0, :, 0], [2.1] * 4) and np.array_equal(r["ts"][0], [0.0] * 4)
    assert r["stats"][0, 0] == 0 and r["stats"][1, 0] > 0
    assert abs(r["ys"][1, -1, 0] - 0.7 / math.e) < 1e-5


def test_reverse_time_equals_negated_field():
    # test_integrate.py:325-413: solving backwards == solving the negated field forwards (ts negated, same ys)
    y0 = np.array([[1.3, -0.4]])
    ts = np.linspace(0.0, 2.0, 9)
    fwd = oracle.solve("forced_osc", y0, 0.0, 2.0, None, solver="tsit5", params=[1.0, 0.0, 0.0], rtol=1e-8, atol=1e-8,
                       save_ts=ts, save_t1=False)
    # reverse: from t=0 down to t=-2 of y' = f  <=>  forward of y' = -f ; for the undamped oscillator -f is f with y1 -> -y1
    rev = oracle.solve("forced_osc", y0 * [1, -1], 0.0, -2.0, None, solver="tsit5", params=[1.0, 0.0, 0.0],
                       rtol=1e-8, atol=1e-8, save_ts=-ts, save_t1=False)
    assert np.array_equal(rev["ts"][0], -ts)
    assert np.allclose(rev["ys"][0] * [1, -1], fwd["ys"][0], rtol=1e-12, atol=1e-12)
    assert np.array_equal(rev["stats"], fwd["stats"])


@pytest.mark.parametrize("solver,order", [("heun", 2), ("midpoint", 2), ("ralston", 2), ("bosh3", 3), ("dopri5", 5),
                                            ("tsit5", 5), ("dopri8", 8), ("euler", 1)])
def test_convergence_order(solver, order):
    # test_integrate.py:144-190 recipe: fixed steps dt = 2^-k, slope of log error within +-0.9 of order
    y0 = np.array([[1.0, 0.5]])
    exact = solve_ivp(lambda t, y: [y[1], -y[0] + 0.7 * np.sin(2 * t)], (0, 2), y0[0], method="DOP853",
                      rtol=1e-13, atol=1e-13).y[:, -1]
    ks = range(2, 7) if order < 8 else range(0, 4)
    errs, dts = [], []
    for k in ks:
        dt = 2.0 ** -k
        r = oracle.solve("forced_osc", y0, 0.0, 2.0, dt, solver=solver, params=[1.0, 0.7, 2.0], controller="constant")
        errs.append(np.linalg.norm(r["ys"][0, 0] - exact)); dts.append(dt)
        assert r["stats"][0, 0] == round(2.0 / dt)
    slope = np.polyfit(np.log(dts), np.log(errs), 1)[0]
    assert abs(slope - order) < 0.9, (solver, slope, errs)


# DETEST-style nonstiff problems (test_detest.py:40-300) through the generic callback field
def _a1(t, y): return [-y[0]]
def _a3(t, y): return [y[0] * np.cos(t)]
def _a4(t, y): return [y[0] / 4 * (1 - y[0] / 20)]
def _b1(t, y): return [2 * (y[0] - y[0] * y[1]), -(y[1] - y[0] * y[1])]     # test_detest.py:66-75 (Lotka-Volterra variant)
def _b4(t, y):
    r = np.sqrt(y[0] ** 2 + y[1] ** 2)
    return [-y[1] - y[0] * y[2] / r, y[0] - y[1] * y[2] / r, y[0] / r]
def _d1(t, y):                                                                # two-body orbit, eccentricity 0.1
    r3 = (y[0] ** 2 + y[1] ** 2) ** 1.5
    return [y[2], y[3], -y[0] / r3, -y[1] / r3]
def _e2(t, y): return [y[1], (1 - y[0] ** 2) * y[1] - y[0]]                   # van der Pol

DETEST = {"A1": (_a1, [1.0]), "A3": (_a3, [1.0]), "A4": (_a4, [1.0]), "B1": (_b1, [1.0, 3.0]),
          "B4": (_b4, [3.0, 0.0, 0.0]), "D1": (_d1, [0.9, 0.0, 0.0, math.sqrt(1.1 / 0.9)]), "E2": (_e2, [2.0, 0.0])}


@pytest.mark.parametrize("prob", list(DETEST))
@pytest.mark.parametrize("solver", ["tsit5", "dopri5", "dopri8"])
def test_detest_vs_scipy_dop853(prob, solver):
    # test_detest.py:440-469: PID(1e-8,1e-8), dt0=None, max_steps=16**4, t in [0,20], DOP853 at 1e-8 -> within 4e-5
    f, y0 = DETEST[prob]
    ref = solve_ivp(f, (0, 20), y0, method="DOP853", rtol=1e-8, atol=1e-8).y[:, -1]
    r = oracle.solve("callback", np.array([y0]), 0.0, 20.0, None, solver=solver, rtol=1e-8, atol=1e-8,
                     max_steps=16 ** 4, callback=f)
    assert r["result"][0] == 0
    assert np.allclose(r["ys"][0, 0], ref, rtol=4e-5, atol=4e-5), (r["ys"][0, 0], ref)


def test_builtin_fields_match_callback_definitions():
    """Each registered field equals its textbook definition evaluated through the callback path, bit for bit."""
    lor = lambda t, y: [10.0 * (y[1] - y[0]), y[0] * (28.0 - y[2]) - y[1], y[0] * y[1] - (8.0 / 3.0) * y[2]]
    lv = lambda t, y: [1.5 * y[0] + (-1.0 * y[0]) * y[1], -3.0 * y[1] + (1.0 * y[0]) * y[1]]
    vdp = lambda t, y: [y[1], 1.5 * (1 - y[0] * y[0]) * y[1] - y[0]]
    for name, params, f, y0, t1 in (("lorenz", [10.0, 28.0, 8.0 / 3.0], lor, [1.0, 2.0, 20.0], 1.0),
                                    ("lotka_volterra", [1.5, -1.0, -3.0, 1.0], lv, [1.0, 1.0], 5.0),
                                    ("vdp", [1.5], vdp, [2.0, 0.0], 3.0)):
        a = oracle.solve(name, np.array([y0]), 0.0, t1, None, solver="dopri5", params=params, rtol=1e-7, atol=1e-7)
        b = oracle.solve("callback", np.array([y0]), 0.0, t1, None, solver="dopri5", rtol=1e-7, atol=1e-7, callback=f)
        assert np.array_equal(a["ys"], b["ys"]) and np.array_equal(a["stats"], b["stats"])


def test_initial_step_is_constant_001():
    """through diffeqsolve, dt0=None means a first trial step of 0.01."""
    r = oracle.solve("lorenz", np.array([[1.0, 2.0, 20.0]]), 0.0, 2.0, None, solver="dopri5", params=[10, 28, 8 / 3],
                     rtol=1e-8, atol=1e-8, trace_traj=0)
    assert r["trace"][0, 0] == 0.0 and r["trace"][0, 1] == 0.01
    h = oracle.solve("lorenz", np.array([[1.0, 2.0, 20.0]]), 0.0, 2.0, None, solver="dopri5", params=[10, 28, 8 / 3],
                     rtol=1e-8, atol=1e-8, trace_traj=0, hairer_initial_step=True)
    assert h["trace"][0, 1] != 0.01 and np.allclose(h["ys"], r["ys"], rtol=1e-6)


def test_pid_rejection_and_clipping_trace():
    """A too-large dt0 is rejected with factor in [0.2, 0.9]; the last step lands exactly on t1 (App. A5)."""
    r = oracle.solve("vdp", np.array([[2.0, 0.0]]), 0.0, 3.0, 1.0, solver="tsit5", params=[5.0], rtol=1e-6, atol=1e-6,
                     trace_traj=0)
    tr = r["trace"]
    assert tr[0, 2] == 0.0                              # first attempt rejected
    w0, w1 = tr[0, 1] - tr[0, 0], tr[1, 1] - tr[1, 0]
    assert tr[1, 0] == tr[0, 0] and 0.2 * w0 - 1e-15 <= w1 <= 0.9 * w0 + 1e-15
    assert tr[-1, 1] == 3.0 and tr[-1, 2] == 1.0
    kept = tr[tr[:, 2] == 1.0]
    assert np.all(kept[1:, 0] == kept[:-1, 1])          # accepted steps tile [t0, t1]
    assert r["stats"][0, 0] == len(tr) and r["stats"][0, 1] == len(kept)


def test_max_steps_reached_and_dtmin():
    r = oracle.solve("lorenz", np.array([[1.0, 2.0, 20.0]]), 0.0, 50.0, None, solver="dopri5", params=[10, 28, 8 / 3],
                     rtol=1e-8, atol=1e-8, max_steps=64)
    assert r["result"][0] == 1 and r["stats"][0, 0] == 64 and r["t_final"][0] < 50.0
    r = oracle.solve("vdp", np.array([[2.0, 0.0]]), 0.0, 3.0, 0.5, solver="tsit5", params=[50.0], rtol=1e-10, atol=1e-10,
                     dtmin=1e-2, force_dtmin=False, max_steps=10000)
    assert r["result"][0] == 2
    r = oracle.solve("vdp", np.array([[2.0, 0.0]]), 0.0, 3.0, 0.5, solver="tsit5", params=[50.0], rtol=1e-10, atol=1e-10,
                     dtmin=1e-2, force_dtmin=True, max_steps=10000)
    assert r["result"][0] == 0 and r["t_final"][0] == 3.0


def test_dense_interpolation_reproduces_knots_and_solution():
    # test_global_interpolation.py:312-390: first point reproduces y0 EXACTLY; values within 1e-6 of exp(-t)
    for solver in ("tsit5", "dopri5", "dopri8", "heun", "bosh3"):
        r = oracle.solve("decay", np.array([[1.0], [0.5]]), 0.0, 1.0, 1e-2 if solver in ("heun", "bosh3") else 0.05,
                         solver=solver, params=[1.0], controller="constant", save_dense=True, max_steps=256)
        tq = np.tile(np.linspace(0, 1, 101), (2, 1))
        ev = oracle.dense_evaluate(solver, r["dense"], tq)
        assert np.array_equal(ev[:, 0, 0], [1.0, 0.5])
        assert np.allclose(ev[:, :, 0], np.array([[1.0], [0.5]]) * np.exp(-tq), atol=1e-5 if solver in ("heun",) else 1e-6)
        out = oracle.dense_evaluate(solver, r["dense"], np.array([[1.5], [-0.1]]))
        assert np.all(np.isnan(out))                      # _nan_if_out_of_bounds


def test_clip_step_size_controller_semantics():
    """test_adaptive_stepsize_controller.py:19-70 recipe: with step_ts the solver lands exactly on those times, with
    jump_ts it steps to prevbefore(t) and resumes from nextafter(t); a discontinuous field is then integrated without
    rejections piling up at the jump."""
    y0 = np.array([[1.0]])
    r = oracle.solve("decay", y0, 0.0, 2.0, None, solver="tsit5", params=[1.0], rtol=1e-6, atol=1e-9,
                     step_ts=[0.5, 1.0, 1.5], save_steps=1, save_t1=False, max_steps=256)
    ts = r["ts"][0][np.isfinite(r["ts"][0])]
    assert all(t in ts for t in (0.5, 1.0, 1.5, 2.0))
    assert abs(r["ys"][0][len(ts) - 1, 0] - math.exp(-2)) < 1e-6
    r = oracle.solve("decay", y0, 0.0, 2.0, None, solver="dopri5", params=[1.0], rtol=1e-6, atol=1e-9, jump_ts=[0.7],
                     save_steps=1, save_t1=False, max_steps=256, trace_traj=0)
    tr = r["trace"]
    i = int(np.argmin(np.abs(tr[:, 1] - 0.7)))
    assert tr[i, 1] == np.nextafter(0.7, 0.0) and tr[i + 1, 0] == np.nextafter(0.7, 1.0)
    # reverse time: step_ts are negated and re-sorted by wrap(direction) (clip.py:232-236)
    r = oracle.solve("decay", y0, 2.0, 0.0, None, solver="tsit5", params=[1.0], rtol=1e-6, atol=1e-9,
                     step_ts=[0.5, 1.0, 1.5], save_steps=1, save_t1=False, max_steps=256)
    ts = r["ts"][0][np.isfinite(r["ts"][0])]
    assert all(t in ts for t in (0.5, 1.0, 1.5, 0.0)) and np.all(np.diff(ts) < 0)
    # a forcing jump: y' = -y + H(t - 1); with jump_ts the solve matches the analytic solution and rejects less
    def f(t, y):
        return [-y[0] + (1.0 if t > 1.0 else 0.0)]
    exact = math.exp(-2) + (1 - math.exp(-1))
    a = oracle.solve("callback", y0, 0.0, 2.0, None, solver="dopri5", rtol=1e-8, atol=1e-10, jump_ts=[1.0], callback=f)
    b = oracle.solve("callback", y0, 0.0, 2.0, None, solver="dopri5", rtol=1e-8, atol=1e-10, callback=f)
    assert abs(a["ys"][0, 0, 0] - exact) < 1e-8
    assert a["stats"][0, 2] < b["stats"][0, 2]


def test_half_solver_semantics():
    """HalfSolver.step (_solver/base.py:312-341) restated by hand for Euler on y' = -lam*y with the default I-controller:
    y1 = two half steps, y1_alt = one full step, y_error = |y1 - y1_alt|, error order = order + 1 = 2 (base.py:296-299),
    and SaveAt(ts) interpolates the FULL step's dense_info (linear between y0 and y1_alt for Euler)."""
    lam, rtol, atol, t1 = 1.3, 1e-3, 1e-6, 2.0
    r = oracle.solve("decay", np.array([[1.0]]), 0.0, t1, 0.1, solver="half:euler", params=[lam], controller="pid",
                     rtol=rtol, atol=atol, save_t1=False, save_steps=1, max_steps=4096)
    n_acc = int(r["stats"][0, 1])
    t, tn, y, ts, ys, acc, rej = 0.0, 0.1, 1.0, [], [], 0, 0
    floor = t1
    for _ in range(100):
        floor = np.nextafter(floor, -np.inf)
    while t < t1:
        h = tn - t
        thalf = t + 0.5 * h
        yh = y + (thalf - t) * (-lam * y)
        y1 = yh + (tn - thalf) * (-lam * yh)
        y1_alt = y + h * (-lam * y)
        err = abs(y1 - y1_alt)
        scaled = abs(err / (atol + max(abs(y), abs(y1)) * rtol))
        keep = scaled < 1
        factor = min(max(0.9 * (1.0 / scaled) ** (1.0 / 2.0), 1.0 if keep else 0.2), 10.0 if keep else 0.9)
        nt0 = tn if keep else t
        nt1 = nt0 + h * factor
        if nt1 > floor:
            nt1 = t1 if keep else nt0 + 0.5 * (t1 - nt0)
        if keep:
            y = y1; acc += 1; ts.append(nt0); ys.append(y)
        else:
            rej += 1
        t, tn = nt0, nt1
    assert (acc, rej) == (n_acc, int(r["stats"][0, 2]))
    assert np.allclose(r["ts"][0, :n_acc], ts, rtol=1e-13, atol=0)
    assert np.allclose(r["ys"][0, :n_acc, 0], ys, rtol=1e-12, atol=0)
    # twice as accurate as the plain solver at the same (constant) step
    plain = oracle.solve("decay", np.array([[1.0]]), 0.0, 1.0, 0.01, solver="euler", params=[1.0], controller="constant")
    half = oracle.solve("decay", np.array([[1.0]]), 0.0, 1.0, 0.01, solver="half:euler", params=[1.0], controller="constant")
    e_plain, e_half = abs(plain["ys"][0, -1, 0] - math.exp(-1)), abs(half["ys"][0, -1, 0] - math.exp(-1))
    assert 1.8 < e_plain / e_half < 2.2
    # FSAL inner solver: the carried derivative goes first half -> second half -> next step
    r5 = oracle.solve("decay", np.array([[1.0]]), 0.0, 2.0, 0.1, solver="half:tsit5", params=[1.0], controller="pid",
                      rtol=1e-9, atol=1e-12)
    assert abs(r5["ys"][0, -1, 0] - math.exp(-2.0)) < 1e-10


def test_dense_derivative_is_slope_of_evaluate():
    """DenseInterpolation.derivative (_global_interpolation.py:357-368): every local interpolant's analytic derivative
    against central differences of evaluate; NaN out of bounds; direction factor on a reversed solve."""
    rng = np.random.default_rng(0)
    y0 = rng.uniform(-1, 1, (5, 2))
    tq = np.tile(np.array([0.31, 0.77, 1.53]), (5, 1))
    h = 1e-6
    for solver in ("tsit5", "dopri5", "dopri8", "heun", "bosh3", "midpoint", "euler"):
        kw = dict(solver=solver, params=[1.0, 0.7, 2.0], save_dense=True, max_steps=4096)
        kw.update(dict(controller="constant") if solver == "euler" else dict(rtol=1e-5, atol=1e-7))
        r = oracle.solve("forced_osc", y0, 0.0, 2.0, 0.0437, **kw)
        assert np.all(r["result"] == 0)
        d = oracle.dense_evaluate(solver, r["dense"], tq, derivative=True)
        fd = (oracle.dense_evaluate(solver, r["dense"], tq + h) - oracle.dense_evaluate(solver, r["dense"], tq - h)) / (2 * h)
        assert np.abs(d - fd).max() < 5e-8, solver
        oob = oracle.dense_evaluate(solver, r["dense"], np.full((5, 1), 2.5), derivative=True)
        if solver == "euler":
            # linear interpolant: the jvp tangent (y1 - y0) / (t1 - t0) does not involve the (NaN) primal, so out of
            # bounds the reference returns the slope of the interval NaN indexes to - the last one - not NaN
            last = oracle.dense_evaluate(solver, r["dense"], np.full((5, 1), 1.999), derivative=True)
            assert np.array_equal(oob, last)
        else:
            assert np.all(np.isnan(oob))
    rr = oracle.solve("forced_osc", y0, 2.0, 0.0, -0.05, solver="tsit5", params=[1.0, 0.7, 2.0], save_dense=True, rtol=1e-8, atol=1e-10)
    d = oracle.dense_evaluate("tsit5", rr["dense"], tq, direction=-1.0, derivative=True)
    fd = (oracle.dense_evaluate("tsit5", rr["dense"], tq + h, direction=-1.0) - oracle.dense_evaluate("tsit5", rr["dense"], tq - h, direction=-1.0)) / (2 * h)
    assert np.abs(d - fd).max() < 5e-8


def test_event_bouncing_ball_docstring():
    """_event.py:74-110: x'' = -8, x(0) = 10, Tsit5, dt0 = 0.1 constant, Event(x, Newton(1e-5, 1e-5)):
    'Event time: 1.58...', 'Velocity at event time: -12.64...' (exact: sqrt(2.5), -8 sqrt(2.5))."""
    vf = lambda t, y: np.array([y[1], -8.0])
    kw = dict(solver="tsit5", callback=vf, controller="constant", event="affine", event_params=[1.0, 0.0, 0.0, 0.0], max_steps=4096)
    r = oracle.solve("callback", np.array([[10.0, 0.0]]), 0.0, 1e3, 0.1, event_root=(1e-5, 1e-5), **kw)
    assert r["result"][0] == 3                                            # RESULTS.event_occurred
    assert abs(r["ts"][0, 0] - math.sqrt(2.5)) < 1e-9 and abs(r["ys"][0, 0, 1] + 8 * math.sqrt(2.5)) < 1e-8
    assert f"{r['ts'][0, 0]:.2f}" == "1.58" and f"{r['ys'][0, 0, 1]:.2f}" == "-12.65"
    # without a root finder the solve stops at the end of the step on which the sign changed
    r0 = oracle.solve("callback", np.array([[10.0, 0.0]]), 0.0, 1e3, 0.1, **kw)
    assert r0["result"][0] == 3 and abs(r0["ts"][0, 0] - 1.6) < 1e-12 and r0["stats"][0, 0] == 16
    # direction=True (upcrossing only) ignores the downward crossing: the ball falls until max_steps
    r1 = oracle.solve("callback", np.array([[10.0, 0.0]]), 0.0, 3.0, 0.1, event_direction=True, **kw)
    assert r1["result"][0] == 0 and abs(r1["ts"][0, 0] - 3.0) < 1e-12
    # SaveAt(ts) + steps: nothing after the event time survives (unsave, _integrate.py:777-806); t1 re-saved (862-872)
    ts = np.linspace(0.0, 3.0, 31)
    r2 = oracle.solve("callback", np.array([[10.0, 0.0]]), 0.0, 3.0, 0.1, event_root=(1e-8, 1e-8), save_ts=ts, **kw)
    n_before = int((ts <= math.sqrt(2.5)).sum())
    assert np.all(np.isfinite(r2["ts"][0, :n_before])) and np.all(np.isinf(r2["ts"][0, n_before + 1:]))
    assert abs(r2["ts"][0, n_before] - math.sqrt(2.5)) < 1e-9             # the t1 slot holds (tfinal, yfinal)
    assert np.allclose(r2["ys"][0, :n_before, 0], 10 - 4 * ts[:n_before] ** 2, atol=1e-9)
    # steady state (boolean): y' = -y reaches |f| < 1e-3 at t = ln(1000); detected at the end of that step
    r3 = oracle.solve("decay", np.array([[1.0]]), 0.0, 100.0, 0.01, solver="tsit5", params=[1.0], rtol=1e-6, atol=1e-9,
                      event="steady_state", event_params=[0.0, 1e-3])
    assert r3["result"][0] == 3 and math.log(1000) <= r3["ts"][0, 0] < math.log(1000) + 1.0
    assert abs(r3["ys"][0, 0, 0]) < 1e-3


def test_store_rejected_steps_revisits_rejected_times():
    """ClipStepSizeController(store_rejected_steps=K), clip.py:292-299, 398-424: the end of every rejected step is pushed
    on a stack and later steps are clipped to it, so each rejected time is eventually the end of an accepted step; a
    stack that overflows ends the solve with RESULTS.max_steps_rejected."""
    y0 = np.array([[1.0, 0.5], [-0.3, 1.2], [2.0, -1.0]])
    kw = dict(solver="bosh3", params=[1.0, 0.7, 2.0], rtol=1e-5, atol=1e-7, save_t1=False, save_steps=1, max_steps=4096)
    for traj in range(3):
        r = oracle.solve("forced_osc", y0, 0.0, 6.0, 1.5, store_rejected_steps=16, trace_traj=traj, **kw)
        assert r["result"][traj] == 0
        tr = r["trace"]
        rej = [row[1] for row in tr if row[2] == 0]
        acc = [row[1] for row in tr if row[2] == 1]
        assert len(rej) >= 1
        assert all(any(x == t for x in acc) for t in rej)          # revisited exactly
    plain = oracle.solve("forced_osc", y0, 0.0, 6.0, 1.5, **kw)
    withstack = oracle.solve("forced_osc", y0, 0.0, 6.0, 1.5, store_rejected_steps=16, **kw)
    assert np.all(withstack["stats"][:, 1] >= plain["stats"][:, 1])  # clipping to old rejected times costs steps
    # a one-slot stack overflows as soon as two rejections are pending
    small = oracle.solve("forced_osc", y0, 0.0, 6.0, 5.0, store_rejected_steps=1, **kw)
    assert np.any(small["result"] == 5)                             # max_steps_rejected


def test_getting_started_ode_printout():
    """docs/usage/getting-started.md:22-36: Dopri5, dy/dt = -y on [0, 3], dt0 = 0.1, PIDController(1e-5, 1e-5),
    SaveAt(ts=[0, 1, 2, 3]) prints ts [0. 1. 2. 3.] and ys [1. 0.368 0.135 0.0498] (fp32 in the docs)."""
    for dtype in (np.float32, np.float64):
        r = oracle.solve("decay", np.ones((1, 1), dtype), 0.0, 3.0, 0.1, solver="dopri5", params=[1.0], dtype=dtype, rtol=1e-5,
                         atol=1e-5, save_ts=np.array([0.0, 1.0, 2.0, 3.0]), save_t1=False)
        assert np.array_equal(r["ts"][0], [0.0, 1.0, 2.0, 3.0])
        assert [f"{v:.3g}" for v in r["ys"][0, :, 0]] == ["1", "0.368", "0.135", "0.0498"]


def test_wide_state_callback_against_scipy_dop853():
    """The oracle at a wide state (Lorenz-96, 40 components - what the warp-per-trajectory kernel is checked against,
    tests/test_user_field.py) agrees with scipy's independent DOP853 at tight tolerance (the test_detest.py:451-459 recipe)."""
    from scipy.integrate import solve_ivp
    D, F = 40, 8.0
    f = lambda t, y: (np.roll(y, -1) - np.roll(y, 2)) * np.roll(y, 1) - y + F      # noqa: E731
    rng = np.random.default_rng(40)
    y0 = F + rng.normal(0, 0.5, (3, D))
    r = oracle.solve("callback", y0, 0.0, 0.5, 0.01, solver="dopri8", rtol=1e-11, atol=1e-11, max_steps=100000, callback=f)
    assert np.all(r["result"] == 0)
    for i in range(3):
        s = solve_ivp(f, (0.0, 0.5), y0[i], method="DOP853", rtol=1e-12, atol=1e-12)
        assert np.abs(r["ys"][i, 0] - s.y[:, -1]).max() < 1e-8 * np.abs(s.y[:, -1]).max()
