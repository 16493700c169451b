"""ctypes front-end of the CPU oracle.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package.  The product (``diffrax_b200``)
never does - it fails loudly when its CUDA library is missing instead of falling
back to anything here.

Parity status: "parity unpinned" against a live Diffrax (jax is not importable in
the authoring container, and the reference ships no golden vectors for this path);
pinned to the reference's own offline anchors in ``tests/test_oracle_*.py``.
"""
from __future__ import annotations

import ctypes as C
import math
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")
_LIB_FMA_PATH = os.path.join(_HERE, "liboracle_fma.so")

F64, F32 = 0, 1
SOLVERS = {"tsit5": 0, "dopri5": 1, "dopri8": 2, "heun": 3, "bosh3": 4, "midpoint": 5,
           "ralston": 6, "euler": 7, "shark": 8}
FIELDS = {"decay": 0, "lotka_volterra": 1, "lorenz": 2, "cr3bp": 3, "mlp": 4, "ou": 5,
          "forced_osc": 6, "vdp": 7, "gbm": 8, "ou_matrix2": 18, "ou_matrix3": 19, "callback": 100}
LEVY = {None: 0, "none": 0, "bi": 1, "brownian_increment": 1, "stla": 2, "space_time": 2}
CTRL_CONSTANT, CTRL_PID = 0, 1

CALLBACK = C.CFUNCTYPE(None, C.c_double, C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_int)


class Desc(C.Structure):
    _fields_ = [
        ("field_id", C.c_int32), ("dim", C.c_int32), ("dtype", C.c_int32), ("solver_id", C.c_int32),
        ("field_params", C.c_void_p), ("n_field_params", C.c_int32),
        ("callback", CALLBACK),
        ("n_traj", C.c_int64),
        ("y0", C.c_void_p),
        ("t0", C.c_double), ("t1", C.c_double),
        ("t0_per_traj", C.c_void_p), ("t1_per_traj", C.c_void_p),
        ("dt0", C.c_double),
        ("controller", C.c_int32),
        ("rtol", C.c_double), ("atol", C.c_double), ("pcoeff", C.c_double), ("icoeff", C.c_double),
        ("dcoeff", C.c_double), ("safety", C.c_double), ("factormin", C.c_double), ("factormax", C.c_double),
        ("dtmin", C.c_double), ("dtmax", C.c_double),
        ("force_dtmin", C.c_int32),
        ("error_order", C.c_double),
        ("hairer_initial_step", C.c_int32),
        ("step_ts", C.c_void_p), ("n_step_ts", C.c_int32), ("jump_ts", C.c_void_p), ("n_jump_ts", C.c_int32), ("store_rejected_steps", C.c_int32),
        ("save_t0", C.c_int32), ("save_t1", C.c_int32), ("save_steps", C.c_int32), ("save_dense", C.c_int32),
        ("save_ts", C.c_void_p), ("n_save_ts", C.c_int32), ("max_steps", C.c_int32),
        ("out_size", C.c_int32),
        ("ts_out", C.c_void_p), ("ys_out", C.c_void_p), ("stats", C.c_void_p), ("result", C.c_void_p),
        ("dense_ts", C.c_void_p), ("dense_y0", C.c_void_p), ("dense_y1", C.c_void_p), ("dense_k", C.c_void_p),
        ("dense_count", C.c_void_p), ("y_final", C.c_void_p), ("t_final", C.c_void_p),
        ("levy_area", C.c_int32), ("bm_keys", C.c_void_p),
        ("bm_t0", C.c_double), ("bm_t1", C.c_double), ("bm_tol", C.c_double),
        ("threefry_partitionable", C.c_int32), ("bm_dim", C.c_int32),
        ("n_events", C.c_int32), ("event_kind", C.c_int32 * 4), ("event_direction", C.c_int32 * 4),
        ("event_root_find", C.c_int32),
        ("event_params", C.c_void_p), ("n_event_params", C.c_int32),
        ("event_rtol", C.c_double), ("event_atol", C.c_double),
        ("state_in", C.c_void_p), ("state_in_flags", C.c_int32), ("state_out", C.c_void_p),
        ("trace_traj", C.c_int64), ("trace", C.c_void_p),
        ("num_threads", C.c_int32),
    ]


def build(force: bool = False) -> str:
    """Compile liboracle.so (and the FMA-contracted variant liboracle_fma.so) with the committed Makefile (gcc only)."""
    srcs = [os.path.join(_HERE, f) for f in ("oracle.c", "oracle_core.inc", "oracle.h", "oracle_tableaux.h", "Makefile")]
    for target in (_LIB_PATH, _LIB_FMA_PATH):
        stale = force or not os.path.exists(target) or any(os.path.getmtime(s) > os.path.getmtime(target) for s in srcs)
        if stale:
            subprocess.run(["make", "-C", _HERE, "-s", "-B", os.path.basename(target)], check=True)
    return _LIB_PATH


_lib = None
_libs = {}
_variant = "strict"


class rounding:
    """``with oracle.rounding("fma"): ...`` runs the oracle built with FMA contraction allowed (liboracle_fma.so)."""

    def __init__(self, variant):
        assert variant in ("strict", "fma")
        self.variant = variant

    def __enter__(self):
        global _variant
        self.prev, _variant = _variant, self.variant

    def __exit__(self, *exc):
        global _variant
        _variant = self.prev


def _load(path):
    L = C.CDLL(path)
    L.orc_solve.argtypes = [C.POINTER(Desc)]
    L.orc_solve.restype = C.c_int
    L.orc_out_size.argtypes = [C.POINTER(Desc)]
    L.orc_out_size.restype = C.c_int
    L.orc_last_error.restype = C.c_char_p
    L.orc_num_stages.argtypes = [C.c_int]
    L.orc_num_stages.restype = C.c_int
    L.orc_hw_threads.restype = C.c_int
    L.orc_threefry2x32.argtypes = [C.c_uint32] * 4 + [C.POINTER(C.c_uint32)] * 2
    L.orc_split.argtypes = [C.c_uint32, C.c_uint32, C.c_int, C.c_int, C.c_void_p]
    L.orc_normal_f64.argtypes = [C.c_uint32, C.c_uint32, C.c_int]
    L.orc_normal_f64.restype = C.c_double
    L.orc_normal_f32.argtypes = [C.c_uint32, C.c_uint32, C.c_int]
    L.orc_normal_f32.restype = C.c_float
    L.orc_erfinv_f64.argtypes = [C.c_double]
    L.orc_erfinv_f64.restype = C.c_double
    L.orc_erfinv_f32.argtypes = [C.c_float]
    L.orc_erfinv_f32.restype = C.c_float
    L.orc_normal_vec_f64.argtypes = [C.c_uint32, C.c_uint32, C.c_int, C.c_int, C.c_int]
    L.orc_normal_vec_f64.restype = C.c_double
    L.orc_normal_vec_f32.argtypes = [C.c_uint32, C.c_uint32, C.c_int, C.c_int, C.c_int]
    L.orc_normal_vec_f32.restype = C.c_float
    L.orc_log1p_f64.argtypes = [C.c_double]
    L.orc_log1p_f64.restype = C.c_double
    L.orc_log1p_f32.argtypes = [C.c_float]
    L.orc_log1p_f32.restype = C.c_float
    L.orc_vbt_evaluate.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int64, C.c_void_p, C.c_double, C.c_double,
                                   C.c_double, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int]
    L.orc_vbt_evaluate.restype = C.c_int
    L.orc_dense_evaluate.argtypes = [C.c_int, C.c_int, C.c_int64, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                     C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, C.c_void_p, C.c_int,
                                     C.c_void_p]
    L.orc_dense_evaluate.restype = C.c_int
    L.orc_dense_derivative.argtypes = L.orc_dense_evaluate.argtypes
    L.orc_dense_derivative.restype = C.c_int
    return L


def lib():
    if _variant not in _libs:
        build()
        _libs[_variant] = _load(_LIB_PATH if _variant == "strict" else _LIB_FMA_PATH)
    return _libs[_variant]


def hw_threads() -> int:
    return lib().orc_hw_threads()


def threefry2x32(k0, k1, x0, x1):
    a, b = C.c_uint32(), C.c_uint32()
    lib().orc_threefry2x32(k0, k1, x0, x1, C.byref(a), C.byref(b))
    return a.value, b.value


def split(key, num, partitionable=True):
    out = np.zeros((num, 2), np.uint32)
    lib().orc_split(int(key[0]), int(key[1]), num, int(partitionable), out.ctypes.data)
    return out


def prng_key(seed: int):
    """jax.random.key(seed) / PRNGKey(seed): words (hi32, lo32)."""
    return np.array([(seed >> 32) & 0xFFFFFFFF, seed & 0xFFFFFFFF], np.uint32)


def normal(key, dtype=np.float64, partitionable=True, shape=()):
    """jax.random.normal(key, shape, dtype) for shape () or (m,)."""
    f64 = np.dtype(dtype) == np.float64
    fn = lib().orc_normal_vec_f64 if f64 else lib().orc_normal_vec_f32
    m = shape[0] if shape else 1
    out = np.array([fn(int(key[0]), int(key[1]), w, m, int(partitionable)) for w in range(m)], dtype)
    return out if shape else out[0]


def log1p(x, dtype=np.float64):
    """The explicitly sequenced log1p of the normal draw (oracle.c), elementwise."""
    f64 = np.dtype(dtype) == np.float64
    fn = lib().orc_log1p_f64 if f64 else lib().orc_log1p_f32
    xs = np.asarray(x, dtype)
    return np.array([fn(float(v)) for v in xs.ravel()], dtype).reshape(xs.shape)


def erfinv(x, dtype=np.float64):
    if np.dtype(dtype) == np.float64:
        return lib().orc_erfinv_f64(float(x))
    return np.float32(lib().orc_erfinv_f32(float(x)))


def _ptr(a):
    return None if a is None else a.ctypes.data


HALF = 0x100  # ORC_HALF: HalfSolver(inner) is spelled "half:<inner>", e.g. "half:euler"


def solver_id(name):
    return (HALF | SOLVERS[name[5:]]) if name.startswith("half:") else SOLVERS[name]


def solve(field, y0, t0, t1, dt0=None, *, solver="dopri5", params=(), dtype=np.float64,
          controller="pid", rtol=1e-3, atol=1e-6, pcoeff=0.0, icoeff=1.0, dcoeff=0.0, safety=0.9,
          factormin=0.2, factormax=10.0, dtmin=None, dtmax=None, force_dtmin=True, error_order=None,
          hairer_initial_step=False, save_t0=False, save_t1=True, save_ts=None, save_steps=0,
          save_dense=False, max_steps=4096, levy_area=None, keys=None, bm_t0=0.0, bm_t1=1.0,
          bm_tol=1e-3, partitionable=True, callback=None, trace_traj=None, num_threads=0,
          t0_per_traj=None, t1_per_traj=None, step_ts=None, jump_ts=None,
          event=None, event_params=(), event_direction=None, event_root=None,
          state_in=None, state_in_flags=7, save_state=False, store_rejected_steps=None, bm_dim=0):
    """Run the oracle on a batch.  Mirrors one vmapped diffeqsolve call of the reference."""
    L = lib()
    dt = np.dtype(dtype)
    y0 = np.ascontiguousarray(y0, dt)
    if y0.ndim == 1:
        y0 = y0[:, None]
    n, d = y0.shape
    D = Desc()
    D.field_id = FIELDS[field] if isinstance(field, str) else int(field)
    D.dim, D.dtype, D.solver_id = d, (F64 if dt == np.float64 else F32), solver_id(solver)
    p = np.ascontiguousarray(params, np.float64).ravel()
    D.field_params, D.n_field_params = _ptr(p), p.size
    cb = None
    if callback is not None:
        def _cb(t, yp, op, dim):
            yv = np.array([yp[i] for i in range(dim)])
            o = callback(t, yv)
            for i in range(dim):
                op[i] = float(o[i])
        cb = CALLBACK(_cb)
        D.callback = cb
        D.field_id = FIELDS["callback"]
        num_threads = 1
    D.n_traj, D.y0 = n, _ptr(y0)
    D.t0, D.t1 = float(t0), float(t1)
    t0a = None if t0_per_traj is None else np.ascontiguousarray(t0_per_traj, dt)
    t1a = None if t1_per_traj is None else np.ascontiguousarray(t1_per_traj, dt)
    D.t0_per_traj, D.t1_per_traj = _ptr(t0a), _ptr(t1a)
    D.dt0 = math.nan if dt0 is None else float(dt0)
    D.controller = CTRL_PID if controller == "pid" else CTRL_CONSTANT
    D.rtol, D.atol, D.pcoeff, D.icoeff, D.dcoeff = rtol, atol, pcoeff, icoeff, dcoeff
    D.safety, D.factormin, D.factormax = safety, factormin, factormax
    D.dtmin = math.nan if dtmin is None else dtmin
    D.dtmax = math.nan if dtmax is None else dtmax
    D.force_dtmin = int(force_dtmin)
    D.error_order = math.nan if error_order is None else float(error_order)
    D.hairer_initial_step = int(hairer_initial_step)
    sta = None if step_ts is None else np.sort(np.ascontiguousarray(step_ts, dt))
    jta = None if jump_ts is None else np.sort(np.ascontiguousarray(jump_ts, dt))
    D.step_ts, D.n_step_ts = _ptr(sta), (0 if sta is None else sta.size)
    D.jump_ts, D.n_jump_ts = _ptr(jta), (0 if jta is None else jta.size)
    D.save_t0, D.save_t1, D.save_steps, D.save_dense = int(save_t0), int(save_t1), int(save_steps), int(save_dense)
    tsa = None if save_ts is None else np.ascontiguousarray(save_ts, dt)
    D.save_ts, D.n_save_ts, D.max_steps = _ptr(tsa), (0 if tsa is None else tsa.size), int(max_steps)
    T = L.orc_out_size(C.byref(D))
    D.out_size = T
    ts_out = np.empty((n, T), dt)
    ys_out = np.empty((n, T, d), dt)
    stats = np.zeros((n, 3), np.int32)
    result = np.zeros(n, np.int32)
    y_final = np.empty((n, d), dt)
    t_final = np.empty(n, dt)
    D.ts_out, D.ys_out, D.stats, D.result = _ptr(ts_out), _ptr(ys_out), _ptr(stats), _ptr(result)
    D.y_final, D.t_final = _ptr(y_final), _ptr(t_final)
    s = L.orc_num_stages(D.solver_id)
    dense = None
    if save_dense:
        dts = np.empty((n, max_steps + 1), dt)
        dy0 = np.empty((n, max_steps, d), dt)
        dy1 = np.empty((n, max_steps, d), dt)
        dk = np.empty((n, max_steps, s, d), dt) if solver.split(":")[-1] not in ("euler", "shark") else None
        dcount = np.zeros(n, np.int32)
        D.dense_ts, D.dense_y0, D.dense_y1, D.dense_k, D.dense_count = _ptr(dts), _ptr(dy0), _ptr(dy1), _ptr(dk), _ptr(dcount)
        dense = dict(ts=dts, y0=dy0, y1=dy1, k=dk, count=dcount)
    D.levy_area = LEVY[levy_area]
    ka = None
    if keys is not None:
        ka = np.ascontiguousarray(keys, np.uint32).reshape(n, 2)
    D.bm_keys = _ptr(ka)
    D.bm_t0, D.bm_t1, D.bm_tol = float(bm_t0), float(bm_t1), float(bm_tol)
    D.threefry_partitionable = int(partitionable)
    D.bm_dim = int(bm_dim)
    if event is not None:
        # event: "affine" (params w[0..d), b, wt) or "steady_state" (params rtol, atol) - or lists of them (then event_params
        # is a list of parameter lists and event_direction a list); event_direction None / True / False;
        # event_root: None or (rtol, atol) of the Newton root finder
        kinds = [event] if isinstance(event, str) else list(event)
        plist = [event_params] if isinstance(event, str) else list(event_params)
        dirs = [event_direction] * len(kinds) if not isinstance(event_direction, (list, tuple)) else list(event_direction)
        D.n_events = len(kinds)
        for i, (k_, d_) in enumerate(zip(kinds, dirs)):
            D.event_kind[i] = {"affine": 1, "steady_state": 2}[k_]
            D.event_direction[i] = 0 if d_ is None else (1 if d_ else 2)
        ev_params = np.ascontiguousarray(np.concatenate([np.asarray(p_, np.float64).ravel() for p_ in plist]), np.float64)
        D.event_params, D.n_event_params = _ptr(ev_params), ev_params.size
        if event_root is not None:
            D.event_root_find, D.event_rtol, D.event_atol = 1, float(event_root[0]), float(event_root[1])
    D.store_rejected_steps = 0 if store_rejected_steps is None else int(store_rejected_steps)
    state_out = None
    if state_in is not None:
        state_in = np.ascontiguousarray(state_in, dt)
        D.state_in, D.state_in_flags = _ptr(state_in), int(state_in_flags)
    if save_state:
        state_out = np.empty((n, 5 + d), dt)
        D.state_out = _ptr(state_out)
    trace = None
    if trace_traj is not None:
        trace = np.full((max_steps, 3), np.nan)
        D.trace_traj, D.trace = int(trace_traj), _ptr(trace)
    D.num_threads = int(num_threads)
    rc = L.orc_solve(C.byref(D))
    if rc != 0:
        raise ValueError(L.orc_last_error().decode())
    out = dict(ts=ts_out, ys=ys_out, stats=stats, result=result, y_final=y_final, t_final=t_final)
    if dense is not None:
        out["dense"] = dense
    if state_out is not None:
        out["state"] = state_out
    if trace is not None:
        out["trace"] = trace[: stats[trace_traj, 0]]
    return out


def vbt_evaluate(keys, ta, tb, *, bm_t0=0.0, bm_t1=1.0, tol=1e-3, levy_area="bi", dtype=np.float64,
                 partitionable=True, shape=()):
    dt = np.dtype(dtype)
    keys = np.ascontiguousarray(keys, np.uint32).reshape(-1, 2)
    n = keys.shape[0]
    ta_a = np.ascontiguousarray(np.broadcast_to(np.asarray(ta, dt), (n,)))
    tb_a = np.ascontiguousarray(np.broadcast_to(np.asarray(tb, dt), (n,)))
    m = int(shape[0]) if shape else 0
    W = np.empty((n, m) if m else n, dt)
    H = np.empty((n, m) if m else n, dt)
    rc = lib().orc_vbt_evaluate(F64 if dt == np.float64 else F32, LEVY[levy_area], int(partitionable), n,
                                keys.ctypes.data, bm_t0, bm_t1, tol, ta_a.ctypes.data, tb_a.ctypes.data, 1,
                                W.ctypes.data, H.ctypes.data, m)
    if rc != 0:
        raise ValueError(lib().orc_last_error().decode())
    return W, H


def dense_evaluate(solver, dense, tq, direction=1.0, derivative=False):
    dts = dense["ts"]
    dt = dts.dtype
    n, msp1 = dts.shape
    d = dense["y0"].shape[-1]
    tq = np.ascontiguousarray(tq, dt).reshape(n, -1)
    nq = tq.shape[1]
    out = np.empty((n, nq, d), dt)
    fn = lib().orc_dense_derivative if derivative else lib().orc_dense_evaluate
    fn(F64 if dt == np.float64 else F32, solver_id(solver), n, d, msp1 - 1, _ptr(dts),
                             _ptr(dense["y0"]), _ptr(dense["y1"]), _ptr(dense["k"]), _ptr(dense["count"]),
                             float(direction), _ptr(tq), nq, _ptr(out))
    return out
