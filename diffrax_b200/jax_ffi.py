"""Reference-side binding of the jax.ffi custom call (what a Diffrax maintainer would add next to `_integrate.py`).

Needs jax (not installable in the authoring container: importing this module without jax raises ImportError; nothing else
in `diffrax_b200` imports it).  Build the handler library first::

    python -m diffrax_b200.build --ffi        # -> diffrax_b200/lib/libdfx_xla_ffi.so (against jax.ffi.include_dir())

`solve_one` is written for ONE trajectory and registered with ``vmap_method="expand_dims"``: under ``jax.vmap`` XLA
passes every operand with a leading batch axis (N for batched operands, 1 for the others) and the handler launches the
ensemble kernel ONCE for the whole batch - the batched design, not one launch per trajectory.  It returns exactly what
the tail of ``diffeqsolve`` (`/root/reference/diffrax/_integrate.py:1478-1543`) needs to assemble a ``Solution``.
"""
from __future__ import annotations

import ctypes
import os

import numpy as np

import jax
import jax.numpy as jnp

from . import _lib

_HERE = os.path.dirname(os.path.abspath(__file__))
FFI_LIB = os.path.join(_HERE, "lib", "libdfx_xla_ffi.so")
_registered = False


def register(path: str = FFI_LIB):
    """Load the handler library and register its four targets on the CUDA platform."""
    global _registered
    if _registered:
        return
    lib = ctypes.CDLL(path)
    for name in ("DfxEnsembleSolveF64", "DfxEnsembleSolveF32", "DfxDenseEvaluateF64", "DfxDenseEvaluateF32"):
        jax.ffi.register_ffi_target(name, jax.ffi.pycapsule(getattr(lib, name)), platform="CUDA")
    _registered = True


def out_size(save_t0, save_t1, n_ts, steps, max_steps):
    """_allocate_output, _integrate.py:1273-1293."""
    out = int(bool(save_t0)) + int(n_ts) + (max_steps // steps if steps else 0)
    if save_t1 and (not steps or max_steps % steps):
        out += 1
    return out


def solve_one(y0, *, field_id, params, solver_id, num_stages, t0, t1, dt0, controller=0, rtol=0.0, atol=0.0, pcoeff=0.0,
              icoeff=1.0, dcoeff=0.0, safety=0.9, factormin=0.2, factormax=10.0, dtmin=None, dtmax=None,
              force_dtmin=True, error_order=None, hairer_initial_step=False, step_ts=None, jump_ts=None,
              store_rejected_steps=0, save_t0=False, save_t1=True, save_ts=None, save_steps=0, save_dense=False,
              max_steps=4096, key=None, levy_area=0, bm_shape=(), bm_t0=0.0, bm_t1=1.0, bm_tol=1e-3, state_in=None,
              state_in_flags=0, save_state=False, event_kind=(), event_direction=(), event_params=(), event_root=None,
              field_weights=None, args=None):
    """One trajectory: y0 [d].  `t0` / `t1` may be Python floats (static attributes) or traced scalars (operands).
    `args` [K]: this trajectory's functor parameters (the `args` of the reference's diffeqsolve; under `jax.vmap` the handler
    receives [N, K]) - for functors compiled with per-trajectory parameters (`fields.CudaField`: pass `field_id=f.field_id_args`
    after `f.ensure_kernel(..., per_traj=True)`).
    Returns (ts [T], ys [T, d], stats [3], result [], y_final [d], t_final [], dense, state_out)."""
    register()
    y0 = jnp.asarray(y0)
    dt = y0.dtype
    d = y0.shape[-1]
    static_t = isinstance(t0, (int, float)) and isinstance(t1, (int, float))
    empty = jnp.zeros((0,), dt)
    t0s = empty if static_t else jnp.asarray(t0, dt).reshape(1)
    t1s = empty if static_t else jnp.asarray(t1, dt).reshape(1)
    ts = empty if save_ts is None else jnp.asarray(save_ts, dt)
    T = out_size(save_t0, save_t1, ts.shape[0], save_steps, max_steps)
    ms = max_steps if save_dense else 0
    nk = num_stages if (save_dense and num_stages > 2) else 0          # Euler / ShARK: two-point interpolants carry no k
    S = jax.ShapeDtypeStruct
    out_types = (S((T,), dt), S((T, d), dt), S((3,), jnp.int32), S((), jnp.int32), S((d,), dt), S((), dt),
                 S((ms + 1 if save_dense else 0,), dt), S((ms, d), dt), S((ms, d), dt), S((ms, nk, d), dt),
                 S(() if save_dense else (0,), jnp.int32), S((5 + d if save_state else 0,), dt))
    name = "DfxEnsembleSolveF64" if dt == jnp.float64 else "DfxEnsembleSolveF32"
    call = jax.ffi.ffi_call(name, out_types, vmap_method="expand_dims")
    nan = float("nan")
    outs = call(
        y0, t0s, t1s, ts,
        empty if step_ts is None else jnp.asarray(step_ts, dt), empty if jump_ts is None else jnp.asarray(jump_ts, dt),
        jnp.zeros((0,), jnp.uint32) if key is None else jax.random.key_data(key).astype(jnp.uint32),
        empty if state_in is None else jnp.asarray(state_in, dt),
        empty if field_weights is None else jnp.asarray(field_weights, dt),
        empty if args is None else jnp.asarray(args, dt),
        field_id=np.int32(field_id), solver_id=np.int32(solver_id), controller=np.int32(controller),
        levy_area=np.int32(levy_area), bm_dim=np.int32(bm_shape[0] if bm_shape else 0),
        t0=float(t0) if static_t else 0.0, t1=float(t1) if static_t else 0.0, dt0=nan if dt0 is None else float(dt0),
        rtol=float(rtol), atol=float(atol), pcoeff=float(pcoeff), icoeff=float(icoeff), dcoeff=float(dcoeff),
        safety=float(safety), factormin=float(factormin), factormax=float(factormax),
        dtmin=nan if dtmin is None else float(dtmin), dtmax=nan if dtmax is None else float(dtmax),
        force_dtmin=np.int32(force_dtmin), error_order=nan if error_order is None else float(error_order),
        hairer_initial_step=np.int32(hairer_initial_step), store_rejected_steps=np.int32(store_rejected_steps or 0),
        save_t0=np.int32(save_t0), save_t1=np.int32(save_t1), save_steps=np.int32(save_steps), save_dense=np.int32(save_dense),
        max_steps=np.int32(max_steps), bm_t0=float(bm_t0), bm_t1=float(bm_t1), bm_tol=float(bm_tol),
        partitionable=np.int32(jax.config.jax_threefry_partitionable), state_in_flags=np.int32(state_in_flags),
        event_kind=np.asarray(event_kind, np.int32), event_direction=np.asarray(event_direction, np.int32),
        event_root_find=np.int32(event_root is not None), event_rtol=float(event_root[0]) if event_root else 0.0,
        event_atol=float(event_root[1]) if event_root else 0.0, event_params=np.asarray(event_params, np.float64),
        params=np.asarray(params, np.float64))
    ts_o, ys_o, stats, result, y_final, t_final, dts, dy0, dy1, dk, dcount, state_out = outs
    dense = dict(ts=dts, y0=dy0, y1=dy1, k=dk, count=dcount) if save_dense else None
    return ts_o, ys_o, stats, result, y_final, t_final, dense, (state_out if save_state else None)


def dense_evaluate(dense, tq, *, solver_id, derivative=False, direction=1.0):
    """DenseInterpolation.evaluate / .derivative (_global_interpolation.py:335-368) on a BATCH: dense["ts"] [N, max_steps+1] ...,
    tq [N, nq] -> [N, nq, d]."""
    register()
    dt = dense["ts"].dtype
    n, nq, d = dense["ts"].shape[0], tq.shape[1], dense["y0"].shape[-1]
    name = "DfxDenseEvaluateF64" if dt == jnp.float64 else "DfxDenseEvaluateF32"
    call = jax.ffi.ffi_call(name, jax.ShapeDtypeStruct((n, nq, d), dt), vmap_method="sequential")
    return call(dense["ts"], dense["y0"], dense["y1"], dense["k"], dense["count"], jnp.asarray(tq, dt),
                solver_id=np.int32(solver_id), derivative=np.int32(derivative), direction=float(direction))


def lorenz_dopri5_example(y0_batch, rtol=1e-8, atol=1e-8):
    """BASELINE config 2 through the custom call: `jax.vmap` over initial conditions -> ONE kernel launch."""
    one = lambda y: solve_one(y, field_id=_lib.FIELD_IDS["lorenz"], params=[10.0, 28.0, 8.0 / 3.0],  # noqa: E731
                              solver_id=_lib.SOLVER_IDS["dopri5"], num_stages=7, t0=0.0, t1=2.0, dt0=None,
                              controller=_lib.CTRL_PID, rtol=rtol, atol=atol)
    return jax.jit(jax.vmap(one))(y0_batch)
