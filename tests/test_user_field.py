"""User-written vector fields (fields.CudaField): what a Python `vector_field(t, y, args)` is to the reference's ODETerm /
ControlTerm (_term.py:174-211, 417-427).  CPU: source generation, nvcc build for sm_100a, launcher registration through
dfx_register_launcher.  GPU: the compiled kernels against the oracle's callback field and against the built-in functors
with the same right-hand sides."""
import math
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))

import torch  # noqa: E402
import diffrax_b200 as dfx  # noqa: E402
from diffrax_b200 import _lib  # noqa: E402
import oracle  # noqa: E402

PENDULUM = "f[0] = y[1]; f[1] = -p[0] * sin(y[0]) - p[1] * y[1];"


def test_user_field_compiles_and_registers():
    pend = dfx.fields.CudaField(2, PENDULUM, params=[9.81, 0.1])
    L = _lib.lib()
    assert pend.field_id >= _lib.FIELD_USER and not pend.is_sde
    pend.ensure_kernel(2, 1, _lib.F64, 0)                      # Dopri5, fp64: nvcc -> lib/user/*.so -> dlopen -> registrar
    assert L.dfx_has_kernel(pend.field_id, 2, 1, _lib.F64, 0) == 1
    assert L.dfx_has_kernel(pend.field_id, 2, 0, _lib.F64, 0) == 0   # only what was asked for
    pend.ensure_kernel(2, 1, _lib.F64, 0)                      # cached
    again = dfx.fields.CudaField(2, PENDULUM, params=[1.0, 0.0])     # same source, other parameters: same kernel
    assert again.field_id == pend.field_id
    other = dfx.fields.CudaField(2, "f[0] = y[1]; f[1] = -y[0];")
    assert other.field_id != pend.field_id
    src = pend.source(0x100 | 3, _lib.F32, 0)
    assert "HalfOf<::dfx::Heun>" in src and "DFX_REGISTER(float" in src and PENDULUM in src


def test_user_field_argument_errors():
    with pytest.raises(ValueError, match="not both"):
        dfx.fields.CudaField(1, "f[0] = -y[0];", diffusion="1.0", noise="gx[0] = x[0];")
    with pytest.raises(ValueError, match="dim"):
        dfx.fields.CudaField(9, "f[0] = 0;")
    gbm = dfx.fields.CudaField(1, "f[0] = p[0] * y[0];", params=[0.1, 0.2], noise="gx[0] = p[1] * y[0] * x[0];")
    with pytest.raises(ValueError, match="additive-noise"):
        gbm.ensure_kernel(1, 8, _lib.F64, 2)                   # ShARK needs additive noise (shark.py:10-30)
    with pytest.raises(ValueError, match="SDE solves need"):
        dfx.fields.CudaField(1, "f[0] = -y[0];").ensure_kernel(1, 3, _lib.F64, 1)
    with pytest.raises(RuntimeError, match="nvcc failed"):
        dfx.fields.CudaField(1, "f[0] = no_such_function(y[0]);").ensure_kernel(1, 1, _lib.F64, 0)


def _np(x):
    return x.detach().cpu().numpy() if isinstance(x, torch.Tensor) else np.asarray(x)


@pytest.mark.gpu
@pytest.mark.parametrize("solver", ["dopri5", "tsit5", "dopri8"])
def test_user_pendulum_against_oracle_callback(dev, solver):
    rng = np.random.default_rng(5)
    y0 = rng.uniform(-1.5, 1.5, (48, 2))
    g, c = 9.81, 0.1
    pend = dfx.fields.CudaField(2, PENDULUM, params=[g, c])
    S = {"dopri5": dfx.Dopri5, "tsit5": dfx.Tsit5, "dopri8": dfx.Dopri8}[solver]
    ts = np.linspace(0.0, 3.0, 7)
    sol = dfx.diffeqsolve(dfx.ODETerm(pend), S(), 0.0, 3.0, None, torch.tensor(y0, device=dev), saveat=dfx.SaveAt(ts=ts),
                          stepsize_controller=dfx.PIDController(rtol=1e-8, atol=1e-8))
    o = oracle.solve("callback", y0, 0.0, 3.0, None, solver=solver, rtol=1e-8, atol=1e-8, save_ts=ts, save_t1=False,
                     callback=lambda t, y: np.array([y[1], -g * math.sin(y[0]) - c * y[1]]))
    st = np.stack([_np(sol.stats[k]) for k in ("num_steps", "num_accepted_steps", "num_rejected_steps")], 1)
    same = np.all(st == o["stats"], axis=1)
    assert same.mean() > 0.95 and np.abs(st[:, 1] - o["stats"][:, 1]).max() <= 1
    err = np.abs(_np(sol.ys)[same] - o["ys"][same]).max() / np.abs(o["ys"]).max()
    assert err < 1e-10, err          # (device sin vs libm sin: last-place differences only)
    assert bool((sol.result == 0).all())


@pytest.mark.gpu
def test_user_lorenz_equals_builtin_bitwise(dev):
    """The same right-hand side as csrc/fields.cuh LorenzField: identical bits, on the device and the host path."""
    rng = np.random.default_rng(1)
    y0 = np.stack([rng.uniform(-15, 15, 3000), rng.uniform(-20, 20, 3000), rng.uniform(5, 45, 3000)], 1)
    user = dfx.fields.CudaField(3, "f[0] = p[0] * (y[1] - y[0]); f[1] = y[0] * (p[1] - y[2]) - y[1]; f[2] = y[0] * y[1] - p[2] * y[2];",
                                params=[10.0, 28.0, 8.0 / 3.0])
    ctrl = dfx.PIDController(rtol=1e-8, atol=1e-8)
    for host in (False, True):
        y = y0 if host else torch.tensor(y0, device=dev)
        a = dfx.diffeqsolve(dfx.ODETerm(user), dfx.Dopri5(), 0.0, 1.0, None, y, stepsize_controller=ctrl)
        b = dfx.diffeqsolve(dfx.ODETerm(dfx.fields.Lorenz()), dfx.Dopri5(), 0.0, 1.0, None, y, stepsize_controller=ctrl)
        assert np.array_equal(_np(a.ys), _np(b.ys)) and np.array_equal(_np(a.stats["num_steps"]), _np(b.stats["num_steps"]))
    # fp32, Tsit5, steps saved, backwards in time
    y32 = torch.tensor(y0[:300].astype(np.float32), device=dev)
    kw = dict(saveat=dfx.SaveAt(steps=True, t0=True), stepsize_controller=dfx.PIDController(rtol=1e-4, atol=1e-5), max_steps=128, throw=False)
    a = dfx.diffeqsolve(dfx.ODETerm(user), dfx.Tsit5(), 0.5, 0.0, None, y32, **kw)
    b = dfx.diffeqsolve(dfx.ODETerm(dfx.fields.Lorenz()), dfx.Tsit5(), 0.5, 0.0, None, y32, **kw)
    assert np.array_equal(_np(a.ys), _np(b.ys)) and np.array_equal(_np(a.ts), _np(b.ts))


@pytest.mark.gpu
def test_user_sde_fields_equal_builtins(dev):
    n = 500
    keys = dfx.random.split(dfx.random.key(7), n)
    y1 = torch.ones(n, 1, device=dev, dtype=torch.float32)
    # additive noise, scalar Brownian motion: OU with ShARK (space-time Levy area), fp32, and HalfSolver(Heun) adaptive, fp64
    uou = dfx.fields.CudaField(1, "f[0] = p[0] * (p[1] - y[0]);", params=[1.0, 0.0, 0.5], diffusion="p[2]")
    bou = dfx.fields.OrnsteinUhlenbeck(1.0, 0.0, 0.5)
    for field_pair in [(uou, bou)]:
        outs = []
        for f in field_pair:
            bm = dfx.VirtualBrownianTree(0.0, 1.0, 2.0 ** -8, (), keys, dfx.SpaceTimeLevyArea)
            terms = dfx.MultiTerm(dfx.ODETerm(f.drift), dfx.ControlTerm(f.diffusion, bm))
            outs.append(dfx.diffeqsolve(terms, dfx.ShARK(), 0.0, 1.0, 2.0 ** -6, y1))
        assert np.array_equal(_np(outs[0].ys), _np(outs[1].ys))
        outs = []
        for f in field_pair:
            bm = dfx.VirtualBrownianTree(0.0, 1.0, 2.0 ** -10, (), keys)
            terms = dfx.MultiTerm(dfx.ODETerm(f.drift), dfx.ControlTerm(f.diffusion, bm))
            outs.append(dfx.diffeqsolve(terms, dfx.HalfSolver(dfx.Heun()), 0.0, 1.0, 0.05, y1.double(),
                                        stepsize_controller=dfx.PIDController(rtol=1e-3, atol=1e-4, pcoeff=0.1, icoeff=0.3), throw=False))
        # adaptive stepping on a Brownian path is chaotic in the step times: the two functors compile to differently contracted
        # FMAs (g = p[2] here, sigma + sigma_t t there), and a last-place difference that flips one accept / reject decision
        # changes the path sampled afterwards - so: (nearly) identical on the majority, and close to the tolerance everywhere
        diff = np.abs(_np(outs[0].ys) - _np(outs[1].ys)).ravel()
        assert (diff < 1e-6).mean() > 0.5, (diff < 1e-6).mean()
        assert diff.max() < 0.05                                            # both are the same strong solution to the tolerance
    # state-dependent noise, scalar Brownian motion, d = 2: geometric Brownian motion with Heun (Stratonovich)
    ugbm = dfx.fields.CudaField(2, "f[0] = p[0] * y[0]; f[1] = p[0] * y[1];", params=[0.1, 0.2],
                                noise="gx[0] = (p[1] * y[0]) * x[0]; gx[1] = (p[1] * y[1]) * x[0];", noise_dim=1)
    bgbm = dfx.fields.GeometricBrownianMotion(0.1, 0.2)
    y2 = torch.tensor(np.random.default_rng(0).uniform(0.5, 2.0, (n, 2)), device=dev)
    outs = []
    for f in (ugbm, bgbm):
        bm = dfx.VirtualBrownianTree(0.0, 1.0, 2.0 ** -9, (), keys)
        terms = dfx.MultiTerm(dfx.ODETerm(f.drift), dfx.ControlTerm(f.diffusion, bm))
        outs.append(dfx.diffeqsolve(terms, dfx.Heun(), 0.0, 1.0, 2.0 ** -7, y2))
    assert np.array_equal(_np(outs[0].ys), _np(outs[1].ys))
    # a [2, 2] diffusion matrix through the general product (noise_dim = 2) against OrnsteinUhlenbeckMatrix
    G = np.array([[0.3, -0.1], [0.2, 0.4]])
    umat = dfx.fields.CudaField(2, "f[0] = p[0] * (p[1] - y[0]); f[1] = p[0] * (p[1] - y[1]);", params=[1.0, 0.2, *G.ravel()],
                                noise="gx[0] = p[2] * x[0] + p[3] * x[1]; gx[1] = p[4] * x[0] + p[5] * x[1];", noise_dim=2)
    bmat = dfx.fields.OrnsteinUhlenbeckMatrix(1.0, 0.2, G)
    outs = []
    for f in (umat, bmat):
        bm = dfx.VirtualBrownianTree(0.0, 1.0, 2.0 ** -9, (2,), keys)
        terms = dfx.MultiTerm(dfx.ODETerm(f.drift), dfx.ControlTerm(f.diffusion, bm))
        outs.append(dfx.diffeqsolve(terms, dfx.Heun(), 0.0, 1.0, 2.0 ** -7, y2))
    assert np.abs(_np(outs[0].ys) - _np(outs[1].ys)).max() < 1e-13
    # the Brownian shape is part of the functor
    with pytest.raises(ValueError, match="Brownian motion with 2 component"):
        bm = dfx.VirtualBrownianTree(0.0, 1.0, 2.0 ** -9, (), keys)
        dfx.diffeqsolve(dfx.MultiTerm(dfx.ODETerm(umat.drift), dfx.ControlTerm(umat.diffusion, bm)), dfx.Heun(), 0.0, 1.0, 2.0 ** -7, y2)


@pytest.mark.gpu
def test_user_field_with_events_dense_and_sharded_entry(dev):
    """The generated kernels are the full ensemble kernels: dense output, an event with root finding, the sharded entry."""
    rng = np.random.default_rng(9)
    y0 = torch.tensor(rng.uniform(0.5, 1.5, (200, 2)), device=dev)
    osc = dfx.fields.CudaField(2, "f[0] = y[1]; f[1] = -p[0] * y[0];", params=[4.0])
    ctrl = dfx.PIDController(rtol=1e-9, atol=1e-9)
    sol = dfx.diffeqsolve(dfx.ODETerm(osc), dfx.Tsit5(), 0.0, 2.0, None, y0, saveat=dfx.SaveAt(dense=True, t1=True), stepsize_controller=ctrl)
    tq = torch.tensor([0.3, 1.1, 1.9], device=dev, dtype=torch.float64)
    ev = _np(sol.interpolation.evaluate(tq))
    a, b = _np(y0)[:, 0:1], _np(y0)[:, 1:2]
    exact = a * np.cos(2 * _np(tq))[None] + b / 2 * np.sin(2 * _np(tq))[None]
    assert np.abs(ev[..., 0] - exact).max() < 1e-7
    # event: x crosses zero downwards, located by Newton on the step's interpolant: x(t_event) = 0 and the exact crossing time
    evs = dfx.diffeqsolve(dfx.ODETerm(osc), dfx.Tsit5(), 0.0, 2.0, None, y0, stepsize_controller=ctrl,
                          event=dfx.Event(dfx.AffineEvent([1.0, 0.0]), dfx.Newton(1e-12, 1e-12), direction=False))
    assert bool(dfx.is_event(evs.result).all())
    assert np.abs(_np(evs.ys)[:, -1, 0]).max() < 1e-8
    t_exact = np.arctan2(_np(y0)[:, 0], -_np(y0)[:, 1] / 2) / 2      # first zero of a cos 2t + (b/2) sin 2t with a, b > 0
    assert np.abs(_np(evs.ts)[:, -1] - t_exact).max() < 1e-7
    sh = dfx.sharded_diffeqsolve(dfx.ODETerm(osc), dfx.Tsit5(), 0.0, 2.0, None, y0, stepsize_controller=ctrl)
    assert torch.equal(sh.y_final, sol.ys[:, -1]) and int(sh.stats["num_failed"]) == 0


def test_builtin_combination_is_instantiated_on_first_use():
    """(field, solver, dtype) pairs that csrc/inst_*.cu does not prebuild are compiled when first asked for - through `prepare`,
    which needs no GPU."""
    from diffrax_b200 import fields
    L = _lib.lib()
    half_mid = dfx.HalfSolver(dfx.Midpoint())
    dfx.prepare(dfx.ODETerm(dfx.fields.VanDerPol(1.0)), half_mid, 0.0, 1.0, 0.1, np.ones((4, 2), np.float32),
                stepsize_controller=dfx.PIDController(1e-3, 1e-5))
    assert L.dfx_has_kernel(_lib.FIELD_IDS["vdp"], 2, half_mid.solver_id, _lib.F32, 0) == 1
    assert fields.ensure_builtin_kernel(dfx.fields.LinearDecay(), 5, 1, _lib.F64, 0, 0)          # LinearDecay with d = 5
    # combinations the C ABI refuses with the reference's message are not compiled
    assert not fields.ensure_builtin_kernel(dfx.fields.GeometricBrownianMotion(), 1, 8, _lib.F64, _lib.LEVY_STLA, 0)
    assert not fields.ensure_builtin_kernel(dfx.fields.Lorenz(), 3, 1, _lib.F64, _lib.LEVY_BI, 0)  # not an SDE functor


@pytest.mark.gpu
def test_on_demand_builtin_kernels_against_the_oracle(dev):
    rng = np.random.default_rng(3)
    # HalfSolver(Midpoint) on van der Pol, fp64, adaptive
    y0 = rng.uniform(0.5, 2.0, (64, 2))
    sol = dfx.diffeqsolve(dfx.ODETerm(dfx.fields.VanDerPol(1.3)), dfx.HalfSolver(dfx.Midpoint()), 0.0, 2.0, 0.05, torch.tensor(y0, device=dev),
                          stepsize_controller=dfx.PIDController(rtol=1e-5, atol=1e-7), saveat=dfx.SaveAt(ts=[0.5, 1.0, 2.0]))
    o = oracle.solve("vdp", y0, 0.0, 2.0, 0.05, solver="half:midpoint", params=[1.3], rtol=1e-5, atol=1e-7, save_ts=[0.5, 1.0, 2.0], save_t1=False)
    st = np.stack([_np(sol.stats[k]) for k in ("num_steps", "num_accepted_steps", "num_rejected_steps")], 1)
    same = np.all(st == o["stats"], axis=1)
    assert same.mean() > 0.9 and np.abs(_np(sol.ys)[same] - o["ys"][same]).max() < 1e-9
    # LinearDecay with d = 5 (prebuilt: 1..3), Dopri5 at fixed steps
    y5 = rng.uniform(0.5, 2.0, (40, 5))
    sol = dfx.diffeqsolve(dfx.ODETerm(dfx.fields.LinearDecay(0.7)), dfx.Dopri5(), 0.0, 1.0, 0.125, torch.tensor(y5, device=dev))
    o = oracle.solve("decay", y5, 0.0, 1.0, 0.125, solver="dopri5", params=[0.7], controller="constant")
    assert np.abs(_np(sol.ys) - o["ys"]).max() < 1e-14 and np.abs(_np(sol.ys)[:, 0] - y5 * np.exp(-0.7)).max() < 1e-7
    # Euler-Maruyama on a 3-dimensional OU process with shape (3,) noise: the Brownian increments are bit-exact, so is the path
    n = 50
    keys = np.random.default_rng(4).integers(0, 2 ** 32, (n, 2), dtype=np.uint64).astype(np.uint32)
    ou = dfx.fields.OrnsteinUhlenbeck(1.2, 0.3, 0.4)
    bm = dfx.VirtualBrownianTree(0.0, 1.0, 2.0 ** -8, (3,), torch.tensor(keys.view(np.int32), device=dev))
    y3 = rng.uniform(0.5, 2.0, (n, 3))
    sol = dfx.diffeqsolve(dfx.MultiTerm(dfx.ODETerm(ou.drift), dfx.ControlTerm(ou.diffusion, bm)), dfx.Euler(), 0.0, 1.0, 2.0 ** -6,
                          torch.tensor(y3, device=dev))
    o = oracle.solve("ou", y3, 0.0, 1.0, 2.0 ** -6, solver="euler", params=[1.2, 0.3, 0.4], controller="constant", levy_area="bi", keys=keys,
                     bm_t0=0.0, bm_t1=1.0, bm_tol=2.0 ** -8, bm_dim=3)
    assert np.abs(_np(sol.ys) - o["ys"]).max() < 1e-13
    # ... and with 6 components (shape (6,)), Heun + space-time Levy area, fp32
    y6 = rng.uniform(0.5, 2.0, (n, 6)).astype(np.float32)
    bm = dfx.VirtualBrownianTree(0.0, 1.0, 2.0 ** -7, (6,), torch.tensor(keys.view(np.int32), device=dev), dfx.SpaceTimeLevyArea)
    sol = dfx.diffeqsolve(dfx.MultiTerm(dfx.ODETerm(ou.drift), dfx.ControlTerm(ou.diffusion, bm)), dfx.Heun(), 0.0, 1.0, 2.0 ** -5,
                          torch.tensor(y6, device=dev))
    o = oracle.solve("ou", y6, 0.0, 1.0, 2.0 ** -5, solver="heun", params=[1.2, 0.3, 0.4], controller="constant", levy_area="stla", keys=keys,
                     bm_t0=0.0, bm_t1=1.0, bm_tol=2.0 ** -7, bm_dim=6, dtype=np.float32)
    assert np.abs(_np(sol.ys) - o["ys"]).max() < 2e-6


def test_user_event_source_and_checks():
    osc = dfx.fields.CudaField(2, "f[0] = y[1]; f[1] = -p[0] * y[0];", params=[4.0], events=["y[0] * y[1]", "y[0] - p[0] * t"])
    plain = dfx.fields.CudaField(2, "f[0] = y[1]; f[1] = -p[0] * y[0];", params=[4.0])
    assert osc.field_id != plain.field_id                                   # the conditions are part of the functor
    src = osc.source(0, _lib.F64, 0)
    assert "kUserEvents = 2" in src and "case 1: return (R)(y[0] - p[0] * t);" in src
    ev = dfx.Event([osc.event(0), osc.event(1)], dfx.Newton(1e-10, 1e-10), direction=[False, None])
    p = dfx.prepare(dfx.ODETerm(osc), dfx.Tsit5(), 0.0, 1.0, 0.1, np.ones((3, 2)), stepsize_controller=dfx.PIDController(1e-6, 1e-6), event=ev)
    assert p.desc.n_events == 2 and list(p.desc.event_kind)[:2] == [_lib.EVENT_USER, _lib.EVENT_USER]
    with pytest.raises(IndexError):
        osc.event(2)
    with pytest.raises(ValueError, match="belongs to the functor"):
        dfx.prepare(dfx.ODETerm(plain), dfx.Tsit5(), 0.0, 1.0, 0.1, np.ones((3, 2)), event=dfx.Event(osc.event(0)))


@pytest.mark.gpu
def test_user_event_conditions(dev):
    """Event(cond_fn) with the user's own condition: x v changes sign first where v does (the turning point, t = phi / 2 for
    x = a cos 2t + (b/2) sin 2t, phi = atan2(b/2, a)); and `y[0]` as a user condition finds the crossing AffineEvent([1, 0]) finds."""
    rng = np.random.default_rng(9)
    y0n = rng.uniform(0.5, 1.5, (200, 2))
    y0 = torch.tensor(y0n, device=dev)
    osc = dfx.fields.CudaField(2, "f[0] = y[1]; f[1] = -p[0] * y[0];", params=[4.0], events=["y[0] * y[1]", "y[0]"])
    ctrl = dfx.PIDController(rtol=1e-10, atol=1e-10)
    a = dfx.diffeqsolve(dfx.ODETerm(osc), dfx.Tsit5(), 0.0, 2.0, None, y0, stepsize_controller=ctrl,
                        event=dfx.Event(osc.event(0), dfx.Newton(1e-12, 1e-12)))
    assert bool(dfx.is_event(a.result).all())
    phi = np.arctan2(y0n[:, 1] / 2, y0n[:, 0])
    assert np.abs(_np(a.ts)[:, -1] - phi / 2).max() < 1e-8
    assert np.abs(_np(a.ys)[:, -1, 1]).max() < 1e-7                                   # v = 0 there
    b = dfx.diffeqsolve(dfx.ODETerm(osc), dfx.Tsit5(), 0.0, 2.0, None, y0, stepsize_controller=ctrl,
                        event=dfx.Event(osc.event(1), dfx.Newton(1e-12, 1e-12), direction=False))
    c = dfx.diffeqsolve(dfx.ODETerm(osc), dfx.Tsit5(), 0.0, 2.0, None, y0, stepsize_controller=ctrl,
                        event=dfx.Event(dfx.AffineEvent([1.0, 0.0]), dfx.Newton(1e-12, 1e-12), direction=False))
    assert torch.equal(b.stats["num_steps"], c.stats["num_steps"]) and torch.equal(b.result, c.result)
    assert np.abs(_np(b.ts) - _np(c.ts)).max() < 1e-10 and np.abs(_np(b.ys) - _np(c.ys)).max() < 1e-9
    # without a root finder: the solve ends at the end of the triggering step, exactly as with the affine condition
    b = dfx.diffeqsolve(dfx.ODETerm(osc), dfx.Tsit5(), 0.0, 2.0, None, y0, stepsize_controller=ctrl, event=dfx.Event(osc.event(1)))
    c = dfx.diffeqsolve(dfx.ODETerm(osc), dfx.Tsit5(), 0.0, 2.0, None, y0, stepsize_controller=ctrl, event=dfx.Event(dfx.AffineEvent([1.0, 0.0])))
    assert torch.equal(b.ts, c.ts) and torch.equal(b.ys, c.ys)


PLAIN = {"DFX_OPT_CHAIN_Y0": 0, "DFX_OPT_LAST_STAGE_F": 0, "DFX_OPT_FAST_PID": 0, "DFX_OPT_ABSMAX_FP64": 0}
LORENZ_SRC = "f[0] = p[0] * (y[1] - y[0]); f[1] = y[0] * (p[1] - y[2]) - y[1]; f[2] = y[0] * y[1] - p[2] * y[2];"


@pytest.mark.gpu
@pytest.mark.parametrize("solver", ["dopri5", "tsit5", "dopri8", "bosh3"])
def test_reference_operation_order_build_stays_under_test(dev, solver):
    """The shipped kernels leave the reference's operation order in four places (stage sums chained onto y0, the last stage's
    k taken from f(y1), the division- / pow-free I-controller, max|y| on the FP64 pipe).  A build of the same kernel with all
    four switched OFF - the reference's own order: vector_tree_dot then y0 + incr (runge_kutta.py:871), pow() and IEEE
    division in the controller (pid.py:476-567) - is compiled here and must (1) match the oracle even more closely and (2)
    agree with the shipped kernel in every accept / reject decision and to the north-star 1e-10 in the states."""
    rng = np.random.default_rng(1)
    n = 4096
    y0n = np.stack([rng.uniform(-15, 15, n), rng.uniform(-20, 20, n), rng.uniform(5, 45, n)], 1)
    y0 = torch.tensor(y0n, device=dev)
    S = {"dopri5": dfx.Dopri5, "tsit5": dfx.Tsit5, "dopri8": dfx.Dopri8, "bosh3": dfx.Bosh3}[solver]
    tol = 1e-6 if solver == "bosh3" else 1e-8
    ctrl = dfx.PIDController(rtol=tol, atol=tol)
    plain = dfx.fields.CudaField(3, LORENZ_SRC, params=[10.0, 28.0, 8.0 / 3.0], defines=PLAIN)
    a = dfx.diffeqsolve(dfx.ODETerm(plain), S(), 0.0, 1.0, 0.01, y0, stepsize_controller=ctrl, max_steps=100000)
    b = dfx.diffeqsolve(dfx.ODETerm(dfx.fields.Lorenz()), S(), 0.0, 1.0, 0.01, y0, stepsize_controller=ctrl, max_steps=100000)
    o = oracle.solve("lorenz", y0n, 0.0, 1.0, 0.01, solver=solver, params=[10.0, 28.0, 8.0 / 3.0], rtol=tol, atol=tol, max_steps=100000)
    sa, sb = _np(a.stats["num_accepted_steps"]), _np(b.stats["num_accepted_steps"])
    ra, rb = _np(a.stats["num_steps"]), _np(b.stats["num_steps"])
    same_ab = (sa == sb) & (ra == rb)
    same_ao = (sa == o["stats"][:, 1]) & (ra == o["stats"][:, 0])
    assert same_ab.mean() > 0.995 and same_ao.mean() > 0.995, (same_ab.mean(), same_ao.mean())
    scale = np.abs(o["ys"]).max()
    e_ao = np.abs(_np(a.ys)[same_ao] - o["ys"][same_ao]).max() / scale
    e_ab = np.abs(_np(a.ys)[same_ab] - _np(b.ys)[same_ab]).max() / scale
    e_bo = np.abs(_np(b.ys)[same_ao & same_ab] - o["ys"][same_ao & same_ab]).max() / scale
    assert e_ao < 1e-10 and e_ab < 1e-10 and e_bo < 1e-10, (e_ao, e_ab, e_bo)


L96 = "fi = (y[(i + 1) % D] - y[(i + D - 2) % D]) * y[(i + D - 1) % D] - y[i] + p[0];"


def _l96(F, D):
    return lambda t, y: (np.roll(y, -1) - np.roll(y, 2)) * np.roll(y, 1) - y + F


def test_wide_field_source_and_checks():
    f = dfx.fields.CudaField(40, L96, params=[8.0], wide=True)
    src = f.source(1, _lib.F64, 0)
    assert "wide_kernel.cuh" in src and "DFX_REGISTER_WIDE(double, UserField, ::dfx::Dopri5)" in src and "kDim = 40" in src
    f.ensure_kernel(40, 1, _lib.F64, 0)
    assert _lib.lib().dfx_has_kernel(f.field_id, 40, 1, _lib.F64, 0) == 1
    with pytest.raises(ValueError, match="wide=True"):
        dfx.fields.CudaField(40, L96, params=[8.0])                      # more than 8 components needs the warp mapping
    with pytest.raises(ValueError, match="ODE functor"):
        dfx.fields.CudaField(40, L96, params=[8.0], wide=True, events=["y[0]"])
    with pytest.raises(ValueError, match="explicit RK tableaux"):
        f.ensure_kernel(40, 0x100 | 3, _lib.F64, 0)
    p = dfx.prepare(dfx.ODETerm(f), dfx.Dopri5(), 0.0, 1.0, None, np.ones((3, 40)), stepsize_controller=dfx.PIDController(1e-6, 1e-6))
    assert p.desc.dim == 40


@pytest.mark.gpu
@pytest.mark.parametrize("D,solver", [(40, "dopri5"), (100, "tsit5"), (33, "dopri8"), (64, "bosh3")])
def test_wide_state_warp_per_trajectory_against_oracle(dev, D, solver):
    """Lorenz-96 with D components, one trajectory per warp (csrc/wide_kernel.cuh), against the oracle's callback field:
    identical accept / reject sequences, saved states (interpolated SaveAt(ts) + t1) to 1e-10."""
    rng = np.random.default_rng(D)
    n = 24
    y0 = 8.0 + rng.normal(0, 0.5, (n, D))
    f = dfx.fields.CudaField(D, L96, params=[8.0], wide=True)
    S = {"dopri5": dfx.Dopri5, "tsit5": dfx.Tsit5, "dopri8": dfx.Dopri8, "bosh3": dfx.Bosh3}[solver]
    tol = 1e-6 if solver == "bosh3" else 1e-9
    ts = [0.1, 0.25, 0.5]
    sol = dfx.diffeqsolve(dfx.ODETerm(f), S(), 0.0, 0.5, 0.01, torch.tensor(y0, device=dev), saveat=dfx.SaveAt(ts=ts, t0=True),
                          stepsize_controller=dfx.PIDController(rtol=tol, atol=tol))
    o = oracle.solve("callback", y0, 0.0, 0.5, 0.01, solver=solver, rtol=tol, atol=tol, save_ts=ts, save_t0=True, save_t1=False, callback=_l96(8.0, D))
    st = np.stack([_np(sol.stats[k]) for k in ("num_steps", "num_accepted_steps", "num_rejected_steps")], 1)
    same = np.all(st == o["stats"], axis=1)
    assert same.mean() >= 0.9, (st[:4], o["stats"][:4])
    assert np.array_equal(_np(sol.ts), o["ts"])
    err = np.abs(_np(sol.ys)[same] - o["ys"][same]).max() / np.abs(o["ys"]).max()
    assert err < 1e-10, err
    assert bool((sol.result == 0).all())


@pytest.mark.gpu
def test_wide_state_modes(dev):
    """Fixed steps + SaveAt(steps) with +inf padding, fp32, backwards in time, per-trajectory t1, the host entry, the sharded
    entry's totals, and a refusal of what the wide kernel does not cover."""
    D, n = 72, 70
    rng = np.random.default_rng(2)
    y0 = rng.uniform(0.5, 2.0, (n, D))
    lam = np.linspace(0.5, 2.0, D)
    dec = dfx.fields.CudaField(D, "fi = -(p[0] + p[1] * (R)i) * y[i];", params=[0.5, 1.5 / (D - 1)], wide=True)
    # constant steps, steps saved: exact t grid, states against the oracle, unfilled slots +inf
    sol = dfx.diffeqsolve(dfx.ODETerm(dec), dfx.Tsit5(), 0.0, 1.0, 0.125, torch.tensor(y0, device=dev), saveat=dfx.SaveAt(steps=True), max_steps=12)
    o = oracle.solve("callback", y0, 0.0, 1.0, 0.125, solver="tsit5", controller="constant", save_steps=1, save_t1=False, max_steps=12,
                     callback=lambda t, y: -lam * y)
    assert np.array_equal(_np(sol.ts), o["ts"]) and np.isinf(_np(sol.ts)[:, 8:]).all() and np.isinf(_np(sol.ys)[:, 8:]).all()
    assert np.abs(_np(sol.ys)[:, :8] - o["ys"][:, :8]).max() < 1e-14
    assert np.abs(_np(sol.ys)[:, 7] - y0 * np.exp(-lam)).max() < 1e-6
    # adaptive, backwards in time, per-trajectory t1, host buffers
    t1 = rng.uniform(-1.0, -0.2, n)
    ctrl = dfx.PIDController(rtol=1e-8, atol=1e-10)
    a = dfx.diffeqsolve(dfx.ODETerm(dec), dfx.Dopri5(), 0.0, t1, None, y0, stepsize_controller=ctrl)                      # numpy in: host entry
    b = dfx.diffeqsolve(dfx.ODETerm(dec), dfx.Dopri5(), 0.0, torch.tensor(t1, device=dev), None, torch.tensor(y0, device=dev), stepsize_controller=ctrl)
    assert np.array_equal(_np(a.ys), _np(b.ys)) and np.array_equal(_np(a.stats["num_steps"]), _np(b.stats["num_steps"]))
    assert np.abs(_np(b.ys)[:, 0] / (y0 * np.exp(-lam[None] * t1[:, None])) - 1).max() < 1e-6
    # fp32
    c = dfx.diffeqsolve(dfx.ODETerm(dec), dfx.Bosh3(), 0.0, 1.0, None, torch.tensor(y0.astype(np.float32), device=dev),
                        stepsize_controller=dfx.PIDController(rtol=1e-4, atol=1e-6))
    assert c.ys.dtype == torch.float32 and np.abs(_np(c.ys)[:, 0] / (y0 * np.exp(-lam)) - 1).max() < 1e-3
    # sharded entry: finals + in-kernel totals
    sh = dfx.sharded_diffeqsolve(dfx.ODETerm(dec), dfx.Dopri5(), 0.0, torch.tensor(t1, device=dev), None, torch.tensor(y0, device=dev), stepsize_controller=ctrl)
    assert torch.equal(sh.y_final, b.ys[:, -1]) and int(sh.stats["num_steps"]) == int(b.stats["num_steps"].sum())
    assert int(sh.stats["max_steps_per_trajectory"]) == int(b.stats["num_steps"].max()) and int(sh.stats["num_failed"]) == 0
    # max_steps reached is reported per trajectory
    z = dfx.diffeqsolve(dfx.ODETerm(dec), dfx.Dopri5(), 0.0, 1.0, None, torch.tensor(y0, device=dev), stepsize_controller=ctrl, max_steps=3, throw=False)
    assert bool((z.result == 1).all())
    with pytest.raises(RuntimeError, match="warp-per-trajectory"):
        dfx.diffeqsolve(dfx.ODETerm(dec), dfx.Dopri5(), 0.0, 1.0, None, torch.tensor(y0, device=dev), stepsize_controller=ctrl,
                        saveat=dfx.SaveAt(dense=True), max_steps=64)


def test_plugin_binds_to_the_loaded_library_instance():
    """A plugin must register into the library instance this process has loaded even when the FILE was rebuilt since (a new
    inode at the same path: the test suite's own `cuda_lib` fixture does that when a header is newer than the objects) - it
    carries no DT_NEEDED entry for libdiffrax_b200.so and resolves the registrar by name (RTLD_GLOBAL)."""
    import shutil
    import subprocess
    L = _lib.lib()
    shutil.copy2(_lib.LIB_PATH, _lib.LIB_PATH + ".swap")
    os.replace(_lib.LIB_PATH + ".swap", _lib.LIB_PATH)                 # same bytes, new inode
    f = dfx.fields.CudaField(2, "f[0] = y[1]; f[1] = -p[0] * y[0] - 0.125 * y[1];", params=[2.0])
    f.ensure_kernel(2, 4, _lib.F32, 0)                                    # Bosh3, fp32
    assert L.dfx_has_kernel(f.field_id, 2, 4, _lib.F32, 0) == 1
    so = [p for p in _lib._plugins if f._hash in p][0]
    needed = subprocess.run(["readelf", "-d", so], capture_output=True, text=True).stdout
    assert "libdiffrax_b200" not in needed


@pytest.mark.gpu
@pytest.mark.parametrize("d,width,dtype", [(2, 32, np.float64), (3, 64, np.float32)])
def test_mlp_field_other_sizes_on_demand(dev, d, width, dtype):
    """eqx.nn.MLP-shaped fields of other sizes than the prebuilt d=4 / width 128 (and fp64): the per-thread functor
    MlpField<d, width>, instantiated on first use, against the oracle's MLP field."""
    mlp = dfx.fields.MLP.init(11, d=d, width=width, dtype=np.dtype(dtype).name)
    assert mlp.field_id >= _lib.FIELD_USER
    rng = np.random.default_rng(d)
    y0 = rng.normal(0, 1, (64, d)).astype(dtype)
    f32 = dtype == np.float32
    rtol, atol = (1e-3, 1e-6) if f32 else (1e-7, 1e-9)
    sol = dfx.diffeqsolve(dfx.ODETerm(mlp), dfx.Tsit5(), 0.0, 2.0, 0.1, torch.tensor(y0, device=dev),
                          stepsize_controller=dfx.PIDController(rtol=rtol, atol=atol), saveat=dfx.SaveAt(ts=[0.5, 1.0, 2.0]))
    o = oracle.solve("mlp", y0, 0.0, 2.0, 0.1, solver="tsit5", params=mlp.oracle_params(), dtype=dtype, rtol=rtol, atol=atol,
                     save_ts=[0.5, 1.0, 2.0], save_t1=False)
    st = np.stack([_np(sol.stats[k]) for k in ("num_steps", "num_accepted_steps", "num_rejected_steps")], 1)
    same = np.all(st == o["stats"], axis=1)
    assert same.mean() > (0.7 if f32 else 0.95)
    err = np.abs(_np(sol.ys)[same] - o["ys"][same]).max() / np.abs(o["ys"]).max()
    assert err < (2e-4 if f32 else 1e-10), err


def test_per_trajectory_args_source_and_checks():
    lor = dfx.fields.CudaField(3, LORENZ_SRC, params=[10.0, 28.0, 8.0 / 3.0])
    src = lor.source(1, _lib.F64, 0, per_traj=True)
    assert "kPerTrajArgs = true" in src and f"kId = {lor.field_id_args}" in src and "kPerTrajArgs" not in lor.source(1, _lib.F64, 0)
    args = np.tile([10.0, 28.0, 8.0 / 3.0], (5, 1))
    p = dfx.prepare(dfx.ODETerm(lor), dfx.Dopri5(), 0.0, 1.0, None, np.ones((5, 3)), args, stepsize_controller=dfx.PIDController(1e-6, 1e-6))
    assert p.desc.field_id == lor.field_id_args and p.desc.n_traj_args == 3 and p.desc.traj_args
    assert _lib.lib().dfx_has_kernel(lor.field_id_args, 3, 1, _lib.F64, 0) == 1
    with pytest.raises(ValueError, match=r"\[N, n_params\]"):
        dfx.prepare(dfx.ODETerm(lor), dfx.Dopri5(), 0.0, 1.0, None, np.ones((5, 3)), args[:, :2], stepsize_controller=dfx.PIDController(1e-6, 1e-6))
    # a built-in functor runs its generated twin under `args`
    q = dfx.prepare(dfx.ODETerm(dfx.fields.Lorenz()), dfx.Dopri5(), 0.0, 1.0, None, np.ones((5, 3)), args, stepsize_controller=dfx.PIDController(1e-6, 1e-6))
    assert q.desc.field_id >= _lib.FIELD_USER and q.desc.n_traj_args == 3
    with pytest.raises(ValueError, match="no per-trajectory-parameter kernel"):
        dfx.prepare(dfx.ODETerm(dfx.fields.MLP.init(1)), dfx.Tsit5(), 0.0, 1.0, None, np.ones((5, 4), np.float32), np.ones((5, 2), np.float32),
                    stepsize_controller=dfx.PIDController(1e-3, 1e-6))


@pytest.mark.gpu
def test_per_trajectory_args_parameter_sweep(dev):
    """diffeqsolve(..., args=[N, n_params]) - the vmapped `args` of the reference, a parameter sweep across the ensemble: every
    trajectory must come out exactly as in a solve of its parameter group with ensemble-wide parameters (device path, the
    chunked host path with its row offsets, the sharded entry); and the same for a wide state."""
    rng = np.random.default_rng(12)
    n = 300_000                                   # host path: more than one chunk
    y0 = np.stack([rng.uniform(-15, 15, n), rng.uniform(-20, 20, n), rng.uniform(5, 45, n)], 1)
    rhos = np.array([14.0, 28.0, 35.0, 99.96])
    grp = rng.integers(0, 4, n)
    args = np.stack([np.full(n, 10.0), rhos[grp], np.full(n, 8.0 / 3.0)], 1)
    lor = dfx.fields.CudaField(3, LORENZ_SRC, params=[10.0, 28.0, 8.0 / 3.0])
    ctrl = dfx.PIDController(rtol=1e-6, atol=1e-6)
    y0d, argsd = torch.tensor(y0, device=dev), torch.tensor(args, device=dev)
    sweep = dfx.diffeqsolve(dfx.ODETerm(lor), dfx.Dopri5(), 0.0, 1.0, None, y0d, argsd, stepsize_controller=ctrl)
    for g, rho in enumerate(rhos):
        m = torch.tensor(grp == g, device=dev)
        one = dfx.diffeqsolve(dfx.ODETerm(dfx.fields.CudaField(3, LORENZ_SRC, params=[10.0, float(rho), 8.0 / 3.0])), dfx.Dopri5(), 0.0, 1.0, None,
                              y0d[m], stepsize_controller=ctrl)
        assert torch.equal(sweep.ys[m], one.ys) and torch.equal(sweep.stats["num_steps"][m], one.stats["num_steps"])
    assert len(set(_np(sweep.stats["num_steps"])[grp == 0].tolist()) ^ set(_np(sweep.stats["num_steps"])[grp == 3].tolist())) > 0  # the sweep matters
    host = dfx.diffeqsolve(dfx.ODETerm(lor), dfx.Dopri5(), 0.0, 1.0, None, y0, args, stepsize_controller=ctrl)            # numpy in: host entry
    assert np.array_equal(_np(host.ys), _np(sweep.ys)) and np.array_equal(_np(host.stats["num_steps"]), _np(sweep.stats["num_steps"]))
    sh = dfx.sharded_diffeqsolve(dfx.ODETerm(lor), dfx.Dopri5(), 0.0, 1.0, None, y0d, argsd, stepsize_controller=ctrl)
    assert torch.equal(sh.y_final, sweep.ys[:, 0]) and int(sh.stats["num_steps"]) == int(sweep.stats["num_steps"].sum())
    # wide state: Lorenz-96 with a per-trajectory forcing F
    D, nw = 40, 600
    l96 = dfx.fields.CudaField(D, L96, params=[8.0], wide=True)
    yw = torch.tensor(8.0 + rng.normal(0, 0.5, (nw, D)), device=dev)
    F = np.where(np.arange(nw) % 2 == 0, 8.0, 4.0)
    a = dfx.diffeqsolve(dfx.ODETerm(l96), dfx.Tsit5(), 0.0, 0.5, None, yw, torch.tensor(F[:, None], device=dev), stepsize_controller=ctrl)
    for val in (8.0, 4.0):
        m = torch.tensor(F == val, device=dev)
        b = dfx.diffeqsolve(dfx.ODETerm(dfx.fields.CudaField(D, L96, params=[val], wide=True)), dfx.Tsit5(), 0.0, 0.5, None, yw[m], stepsize_controller=ctrl)
        assert torch.equal(a.ys[m], b.ys)


@pytest.mark.gpu
def test_builtin_functors_take_per_trajectory_args(dev):
    """`args` on a built-in functor runs its generated twin: with every row equal to the functor's own parameters the result is,
    bit for bit, the built-in kernel's (so the twins ARE the functors of csrc/fields.cuh); with a real sweep, each row's own."""
    rng = np.random.default_rng(21)
    n = 500
    ctrl = dfx.PIDController(rtol=1e-7, atol=1e-9)
    cases = [(dfx.fields.LinearDecay(0.7), 3, 0.0), (dfx.fields.LotkaVolterra(), 2, 0.0), (dfx.fields.Lorenz(), 3, 0.0),
             (dfx.fields.CR3BP(), 4, 0.0), (dfx.fields.ForcedOscillator(1.0, 0.7, 2.0), 2, 0.0), (dfx.fields.VanDerPol(1.5), 2, 0.0)]
    for f, d, _ in cases:
        y0 = torch.tensor(rng.uniform(0.5, 1.5, (n, d)), device=dev)
        if isinstance(f, dfx.fields.CR3BP):
            y0 = torch.tensor(np.array([0.994, 0.0, 0.0, -2.00158510637908252]) + 1e-3 * rng.standard_normal((n, 4)), device=dev)
        same = torch.tensor(np.tile(f.params(), (n, 1)), device=dev)
        a = dfx.diffeqsolve(dfx.ODETerm(f), dfx.Tsit5(), 0.0, 1.0, None, y0, same, stepsize_controller=ctrl, throw=False)
        b = dfx.diffeqsolve(dfx.ODETerm(f), dfx.Tsit5(), 0.0, 1.0, None, y0, stepsize_controller=ctrl, throw=False)
        assert torch.equal(a.ys, b.ys) and torch.equal(a.stats["num_steps"], b.stats["num_steps"]), type(f).__name__
        assert torch.equal(a.result, b.result) and int((b.result == 0).sum()) > n // 2
    # a sweep over the van der Pol stiffness: row i is the solve with mu_i
    mus = np.where(np.arange(n) % 2 == 0, 0.5, 3.0)
    y0 = torch.tensor(rng.uniform(0.5, 1.5, (n, 2)), device=dev)
    sw = dfx.diffeqsolve(dfx.ODETerm(dfx.fields.VanDerPol()), dfx.Tsit5(), 0.0, 2.0, None, y0, torch.tensor(mus[:, None], device=dev), stepsize_controller=ctrl,
                         throw=False)
    for mu in (0.5, 3.0):
        m = torch.tensor(mus == mu, device=dev)
        one = dfx.diffeqsolve(dfx.ODETerm(dfx.fields.VanDerPol(mu)), dfx.Tsit5(), 0.0, 2.0, None, y0[m], stepsize_controller=ctrl, throw=False)
        assert torch.equal(sw.ys[m], one.ys)
    # SDE: OU with a per-path mean, fixed steps
    keys = dfx.random.split(dfx.random.key(5), n)
    ou = dfx.fields.OrnsteinUhlenbeck(1.0, 0.0, 0.5)
    mk = lambda f: dfx.MultiTerm(dfx.ODETerm(f.drift), dfx.ControlTerm(f.diffusion, dfx.VirtualBrownianTree(0.0, 1.0, 2.0 ** -8, (), keys)))  # noqa: E731
    y1 = torch.ones(n, 1, device=dev, dtype=torch.float64)
    pa = np.tile([1.0, 0.0, 0.5, 0.0], (n, 1)); pa[:, 1] = np.where(np.arange(n) % 2 == 0, 0.0, 2.0)
    sw = dfx.diffeqsolve(mk(ou), dfx.Heun(), 0.0, 1.0, 2.0 ** -6, y1, torch.tensor(pa, device=dev))
    base = dfx.diffeqsolve(mk(ou), dfx.Heun(), 0.0, 1.0, 2.0 ** -6, y1)
    even = torch.arange(n, device=dev) % 2 == 0
    assert torch.equal(sw.ys[even], base.ys[even]) and not torch.equal(sw.ys[~even], base.ys[~even])
