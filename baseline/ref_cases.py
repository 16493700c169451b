"""Oracle-style case dictionaries -> live ``diffrax.diffeqsolve`` calls under ``jax.vmap``.

A case is the dict ``tests/golden/make_golden.py`` / ``bench.workload`` build: ``field``, ``params``, ``solver``, ``y0``,
``t0``, ``t1``, ``dt0`` plus the oracle's keyword arguments.  Every registered device functor of ``csrc/fields.cuh`` has
its JAX twin here (same expression order), so the SAME seeded inputs go through the reference's own code path:
``diffeqsolve`` (``/root/reference/diffrax/_integrate.py:888``) vmapped over ``y0`` (and the Brownian keys) as in
``/root/reference/test/helpers.py:136-186`` and ``test/test_vmap.py:28-39``.

Nothing here is importable without jax; ``baseline.probe()`` gates every use.
"""
from __future__ import annotations

import numpy as np


def _field(jnp, name, p):
    """f(t, y, args) for ODETerm - twins of csrc/fields.cuh / oracle_core.inc field_eval."""
    if name == "decay":
        return lambda t, y, args: -p[0] * y
    if name == "lotka_volterra":      # benchmarks/lotka_volterra.py:13-20
        a, b, c, d = p[:4]
        return lambda t, y, args: jnp.stack([a * y[0] + (b * y[0]) * y[1], c * y[1] + (d * y[0]) * y[1]])
    if name == "lorenz":
        s, r, b = p[:3]
        return lambda t, y, args: jnp.stack([s * (y[1] - y[0]), y[0] * (r - y[2]) - y[1], y[0] * y[1] - b * y[2]])
    if name == "cr3bp":
        mu = p[0]
        mup = 1.0 - mu

        def f(t, y, args):
            x, yy, vx, vy = y[0], y[1], y[2], y[3]
            dx1, dx2 = x + mu, x - mup
            r1s, r2s = dx1 * dx1 + yy * yy, dx2 * dx2 + yy * yy
            w1, w2 = mup / (r1s * jnp.sqrt(r1s)), mu / (r2s * jnp.sqrt(r2s))
            return jnp.stack([vx, vy, x + 2.0 * vy - w1 * dx1 - w2 * dx2, yy - 2.0 * vx - (w1 + w2) * yy])
        return f
    if name == "forced_osc":
        w0sq, amp, w = p[:3]
        return lambda t, y, args: jnp.stack([y[1], -w0sq * y[0] + amp * jnp.sin(w * t)])
    if name == "vdp":
        mu = p[0]
        return lambda t, y, args: jnp.stack([y[1], mu * (1.0 - y[0] * y[0]) * y[1] - y[0]])
    if name == "ou":
        theta, mu = p[0], p[1]
        return lambda t, y, args: theta * (mu - y)
    if name == "mlp":                 # eqx.nn.MLP(d -> W -> W -> d, softplus, final tanh), neural_ode.ipynb cell 5
        import jax
        width, depth = int(p[0]), int(p[1])
        q = np.asarray(p[2:], np.float64)
        layers, nin, off = [], None, 0
        d = (q.size - (width * width + width) * (depth - 1) - width) // (2 * width + 1)
        nin = d
        for L in range(depth + 1):
            nout = d if L == depth else width
            W = q[off: off + nout * nin].reshape(nout, nin); off += nout * nin
            b = q[off: off + nout]; off += nout
            layers.append((W, b)); nin = nout

        def f(t, y, args):
            h = y
            for i, (W, b) in enumerate(layers):
                h = jnp.asarray(W, y.dtype) @ h + jnp.asarray(b, y.dtype)
                h = jnp.tanh(h) if i == len(layers) - 1 else jax.nn.softplus(h)
            return h
        return f
    raise ValueError(f"no JAX twin for field {name!r}")


def _solver(dfx, name):
    if name.startswith("half:"):
        return dfx.HalfSolver(_solver(dfx, name[5:]))
    return {"tsit5": dfx.Tsit5, "dopri5": dfx.Dopri5, "dopri8": dfx.Dopri8, "heun": dfx.Heun, "bosh3": dfx.Bosh3,
            "midpoint": dfx.Midpoint, "ralston": dfx.Ralston, "euler": dfx.Euler, "shark": dfx.ShARK}[name]()


def _build(jax, dfx, case):
    import jax.numpy as jnp
    import jax.random as jr
    kw = dict(case)
    dtype = np.dtype(kw.get("dtype", np.float64))
    y0 = jnp.asarray(np.asarray(kw["y0"], dtype))
    if y0.ndim == 1:
        y0 = y0[:, None]
    p = [float(v) for v in np.asarray(kw["params"], np.float64).ravel()]
    f = _field(jnp, kw["field"], p)
    solver = _solver(dfx, kw["solver"])
    t0, t1, dt0 = kw["t0"], kw["t1"], kw["dt0"]
    if kw.get("controller", "pid") == "constant":
        ctrl = dfx.ConstantStepSize()
    else:
        ctrl = dfx.PIDController(rtol=kw["rtol"], atol=kw["atol"], pcoeff=kw.get("pcoeff", 0.0), icoeff=kw.get("icoeff", 1.0),
                                 dcoeff=kw.get("dcoeff", 0.0), dtmin=kw.get("dtmin"), dtmax=kw.get("dtmax"),
                                 force_dtmin=kw.get("force_dtmin", True))
    if kw.get("step_ts") is not None or kw.get("jump_ts") is not None:
        ctrl = dfx.ClipStepSizeController(ctrl, step_ts=kw.get("step_ts"), jump_ts=kw.get("jump_ts"),
                                          store_rejected_steps=kw.get("store_rejected_steps"))
    ts = kw.get("save_ts")
    steps = kw.get("save_steps", 0)
    saveat = dfx.SaveAt(t0=kw.get("save_t0", False), t1=kw.get("save_t1", True),
                        ts=None if ts is None else jnp.asarray(np.asarray(ts, dtype)),
                        steps=bool(steps) if steps in (0, 1) else steps, dense=kw.get("save_dense", False))
    max_steps = kw.get("max_steps", 4096)
    levy = kw.get("levy_area")
    if levy:
        jax.config.update("jax_threefry_partitionable", bool(kw.get("partitionable", True)))
        keys = jr.wrap_key_data(jnp.asarray(np.asarray(kw["keys"], np.uint32)))
        la = dfx.BrownianIncrement if levy in ("bi", "brownian_increment") else dfx.SpaceTimeLevyArea
        shape = (int(kw["bm_dim"]),) if kw.get("bm_dim") else ()
        sigma = p[2]
        sigma_t = p[3] if len(p) > 3 else 0.0
        struct = jax.ShapeDtypeStruct(shape, y0.dtype)
        weak = lambda v: jnp.asarray(v, y0.dtype)  # noqa: E731  (python floats are weakly typed: rounded once to the state dtype)

        def one(y, key):
            bm = dfx.VirtualBrownianTree(kw.get("bm_t0", 0.0), kw.get("bm_t1", 1.0), kw["bm_tol"], struct, key, levy_area=la)
            if shape:   # diagonal diffusion driven by an (m,) Brownian motion
                g = lambda t, y_, args: dfx_lineax_diag(jnp, (weak(sigma) + weak(sigma_t) * t) * jnp.ones_like(y_))  # noqa: E731
            else:       # scalar noise into every component
                g = lambda t, y_, args: (weak(sigma) + weak(sigma_t) * t) * jnp.ones_like(y_)  # noqa: E731
            terms = dfx.MultiTerm(dfx.ODETerm(f), dfx.ControlTerm(g, bm))
            return dfx.diffeqsolve(terms, solver, t0, t1, dt0, y, saveat=saveat, stepsize_controller=ctrl,
                                   max_steps=max_steps, throw=False)
        args = (y0, keys)
    else:
        def one(y):
            return dfx.diffeqsolve(dfx.ODETerm(f), solver, t0, t1, dt0, y, saveat=saveat, stepsize_controller=ctrl,
                                   max_steps=max_steps, throw=False)
        args = (y0,)
    return one, args


def dfx_lineax_diag(jnp, v):
    """Diagonal diffusion as a lineax operator (the documented way to get an elementwise product out of ControlTerm)."""
    import lineax as lx
    return lx.DiagonalLinearOperator(v)


def _pack(sol, dense):
    out = {"ts": np.asarray(sol.ts), "ys": np.asarray(sol.ys),
           "stats": np.stack([np.asarray(sol.stats[k]) for k in ("num_steps", "num_accepted_steps", "num_rejected_steps")], 1),
           "result": np.asarray(sol.result._value)}
    return out


def compiled(jax, dfx, case):
    one, args = _build(jax, dfx, case)
    fn = jax.jit(jax.vmap(one))
    return fn, args


def solve(jax, dfx, case, *, jit=True):
    one, args = _build(jax, dfx, case)
    fn = jax.vmap(one)
    if jit:
        fn = jax.jit(fn)
    sol = fn(*args)
    jax.block_until_ready(sol.ys if sol.ys is not None else sol.stats["num_steps"])
    return _pack(sol, case.get("save_dense", False))
