#!/usr/bin/env python
"""Writes tests/golden/ensemble_golden.npz: small seeded inputs and the oracle's outputs for them.

The reference (Diffrax) cannot be imported in the authoring container (no jax), so these vectors
come from the CPU oracle, which tests/test_oracle_*.py pin to the reference's own offline anchors.
They serve two purposes: (1) the oracle is regression-pinned to them (test_golden_oracle), and
(2) the GPU parity tests compare the CUDA path against them without re-running anything.
Wherever a live Diffrax is importable, `python baseline/gen_golden.py` runs the SAME cases through the reference itself
and writes tests/golden/diffrax_golden.npz, which tests/test_live_reference.py then pins the oracle to.
Usage:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import oracle  # noqa: E402

CASES = {}


def case(name, **kw):
    CASES[name] = kw


rng = np.random.default_rng(2024)
n = 96
case("c1_lv_tsit5_ts", field="lotka_volterra", params=[1.5, -1.0, -3.0, 1.0], solver="tsit5", y0=rng.uniform(0.5, 2, (n, 2)),
     t0=0.0, t1=10.0, dt0=None, rtol=1e-6, atol=1e-6, save_t1=False, save_ts=np.linspace(0, 10, 100))
case("c2_lorenz_dopri5_t1", field="lorenz", params=[10.0, 28.0, 8.0 / 3.0], solver="dopri5",
     y0=np.stack([rng.uniform(-15, 15, n), rng.uniform(-20, 20, n), rng.uniform(5, 45, n)], 1),
     t0=0.0, t1=2.0, dt0=None, rtol=1e-8, atol=1e-8)
y_ar = np.array([0.994, 0.0, 0.0, -2.00158510637908252]) + 1e-4 * rng.standard_normal((32, 4))
case("c3_cr3bp_dopri8_dense", field="cr3bp", params=[0.012277471], solver="dopri8", y0=y_ar, t0=0.0, t1=17.0652165601579625 / 4,
     dt0=None, rtol=1e-12, atol=1e-12, save_dense=True, max_steps=256)
case("steps_t0_t1_bosh3", field="vdp", params=[1.5], solver="bosh3", y0=rng.uniform(-2, 2, (n, 2)), t0=0.0, t1=3.0, dt0=0.05,
     rtol=1e-5, atol=1e-7, save_t0=True, save_t1=True, save_steps=3, max_steps=1000)
case("reverse_time_tsit5", field="forced_osc", params=[1.0, 0.7, 2.0], solver="tsit5", y0=rng.uniform(-2, 2, (n, 2)), t0=1.0,
     t1=-1.5, dt0=None, rtol=1e-7, atol=1e-9, save_ts=np.linspace(1.0, -1.5, 11), save_t1=True)
case("pid_coeffs_dtmax_heun", field="decay", params=[0.7], solver="heun", y0=rng.uniform(0.5, 2, (n, 3)), t0=0.0, t1=2.0,
     dt0=None, rtol=1e-4, atol=1e-6, pcoeff=0.3, icoeff=0.4, dcoeff=0.1, dtmax=0.05, dtmin=1e-4)
case("f32_tsit5", field="lorenz", params=[10.0, 28.0, 8.0 / 3.0], solver="tsit5", dtype=np.float32,
     y0=np.stack([rng.uniform(-15, 15, n), rng.uniform(-20, 20, n), rng.uniform(5, 45, n)], 1).astype(np.float32),
     t0=0.0, t1=1.0, dt0=None, rtol=1e-6, atol=1e-6)
import diffrax_b200 as _dfx  # noqa: E402  (host-side weight initialiser only; no CUDA needed)
_mlp = _dfx.fields.MLP.init(3, d=4, width=128)
case("c4_mlp_tsit5_f32", field="mlp", params=_mlp.oracle_params(), solver="tsit5", dtype=np.float32,
     y0=rng.standard_normal((n, 4)).astype(np.float32), t0=0.0, t1=10.0, dt0=None, rtol=1e-3, atol=1e-6)
keys = oracle.split(oracle.prng_key(2024), n)
for dt_, tag in ((np.float64, "f64"), (np.float32, "f32")):
    case(f"c5_ou_heun_{tag}", field="ou", params=[1.0, 0.0, 0.5], solver="heun", dtype=dt_, y0=np.ones((n, 1), dt_), t0=0.0, t1=1.0,
         dt0=2.0 ** -6, controller="constant", levy_area="bi", keys=keys, bm_tol=2.0 ** -8)
    case(f"c5_ou_shark_{tag}", field="ou", params=[1.0, 0.0, 0.5], solver="shark", dtype=dt_, y0=np.ones((n, 1), dt_), t0=0.0, t1=1.0,
         dt0=2.0 ** -6, controller="constant", levy_area="stla", keys=keys, bm_tol=2.0 ** -8)
case("ou_heun_adaptive_f64", field="ou", params=[1.0, 0.0, 0.5], solver="heun", y0=np.ones((n, 1)), t0=0.0, t1=1.0, dt0=0.1,
     rtol=0.0, atol=1e-2, dtmin=2.0 ** -7, levy_area="bi", keys=keys, bm_tol=2.0 ** -9, pcoeff=0.1, icoeff=0.3)


# Oracle-only regression pins for the SURVEY §8f features (the GPU parity tests build their own inputs for these)
CASES_EXTRA = {}
rng2 = np.random.default_rng(77)
m = 48
yo = rng2.uniform(-2, 2, (m, 2))
CASES_EXTRA["half_heun_ode"] = dict(field="forced_osc", params=[1.0, 0.7, 2.0], solver="half:heun", y0=yo, t0=0.0, t1=3.0, dt0=0.3,
                                    rtol=1e-6, atol=1e-8, save_ts=np.linspace(0.0, 3.0, 7), save_t1=False, max_steps=20000)
CASES_EXTRA["half_shark_sde_fixed"] = dict(field="ou", params=[1.0, 0.0, 0.5], solver="half:shark", y0=np.ones((m, 1)), t0=0.0, t1=1.0,
                                           dt0=2.0 ** -5, controller="constant", levy_area="stla", keys=keys[:m], bm_tol=2.0 ** -9)
CASES_EXTRA["event_newton_tsit5"] = dict(field="forced_osc", params=[1.0, 0.7, 2.0], solver="tsit5", y0=yo, t0=0.0, t1=3.0, dt0=0.3,
                                         rtol=1e-9, atol=1e-11, save_t0=True, save_ts=np.linspace(0.25, 3.0, 12), save_t1=True,
                                         event=["affine", "affine"], event_params=[[1.0, 0.0, -0.3, 0.0], [0.0, 1.0, 2.5, 0.0]],
                                         event_direction=[None, False], event_root=(1e-10, 1e-12))
CASES_EXTRA["clip_steps_jumps_rejected"] = dict(field="forced_osc", params=[1.0, 0.7, 2.0], solver="dopri5", y0=yo, t0=0.0, t1=3.0, dt0=1.0,
                                                rtol=1e-7, atol=1e-9, step_ts=np.array([0.5, 1.75]), jump_ts=np.array([1.0, 2.5]),
                                                store_rejected_steps=8, save_steps=1, save_t1=True, max_steps=512)
CASES_EXTRA["steady_state_decay"] = dict(field="decay", params=[1.0], solver="bosh3", y0=rng2.uniform(0.5, 2.0, (m, 2)), t0=0.0, t1=50.0,
                                         dt0=0.01, rtol=1e-6, atol=1e-4, event="steady_state", event_params=[1e-6, 1e-4])


def main():
    out = {}
    for name, kw in list(CASES.items()) + list(CASES_EXTRA.items()):
        kw = dict(kw)
        field = kw.pop("field")
        y0, t0, t1, dt0 = kw.pop("y0"), kw.pop("t0"), kw.pop("t1"), kw.pop("dt0")
        r = oracle.solve(field, y0, t0, t1, dt0, **kw)
        out[f"{name}/ys"] = r["ys"]; out[f"{name}/ts"] = r["ts"]; out[f"{name}/stats"] = r["stats"]
        out[f"{name}/result"] = r["result"]; out[f"{name}/y_final"] = r["y_final"]
        if "dense" in r:
            out[f"{name}/dense_ts"] = r["dense"]["ts"]; out[f"{name}/dense_count"] = r["dense"]["count"]
            tq = np.tile(np.linspace(t0, t1, 33), (y0.shape[0], 1))
            out[f"{name}/dense_tq"] = tq
            out[f"{name}/dense_eval"] = oracle.dense_evaluate(kw["solver"], r["dense"], tq)
            out[f"{name}/dense_deriv"] = oracle.dense_evaluate(kw["solver"], r["dense"], tq, derivative=True)
        print(name, r["stats"][:3].tolist(), "failed:", int((r["result"] != 0).sum()))
    # PRNG / Brownian golden words
    out["prng/keys"] = keys
    for part in (1, 0):
        out[f"prng/split3_part{part}"] = np.stack([oracle.split(k, 3, bool(part)) for k in keys[:16]])
        for dt_, tag in ((np.float64, "f64"), (np.float32, "f32")):
            out[f"prng/normal_{tag}_part{part}"] = np.array([oracle.normal(k, dt_, bool(part)) for k in keys], dt_)
    for lv in ("bi", "stla"):
        for dt_, tag in ((np.float64, "f64"), (np.float32, "f32")):
            W, H = oracle.vbt_evaluate(keys, 0.3, 0.7, tol=2.0 ** -8, levy_area=lv, dtype=dt_)
            out[f"vbt/{lv}_{tag}_W"] = W; out[f"vbt/{lv}_{tag}_H"] = H
    np.savez_compressed(os.path.join(HERE, "ensemble_golden.npz"), **out)
    print("wrote", os.path.join(HERE, "ensemble_golden.npz"), os.path.getsize(os.path.join(HERE, "ensemble_golden.npz")), "bytes")


if __name__ == "__main__":
    main()
