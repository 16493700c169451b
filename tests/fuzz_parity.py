"""Randomised differential check: the CUDA path (through the public call and the C ABI) against the oracle on random
combinations of field x solver x controller x SaveAt x dtype x direction x per-trajectory t1 x max_steps that no
hand-written test spells out.  Test infrastructure (imports oracle/): run on a GPU box,

    python tests/fuzz_parity.py --cases 400 --seed 0

Per case: result codes and step statistics must agree on (almost) every trajectory; on the trajectories whose step
sequences are identical the saved times / states must agree to 1e-9 (fp64) / 2e-4 (fp32).  Exit code 1 on any failure."""
import argparse
import os
import sys
import traceback

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import diffrax_b200 as dfx  # noqa: E402
import oracle  # noqa: E402

ODE_FIELDS = {
    "decay": (lambda r: [float(r.uniform(0.2, 2.0))], lambda r: int(r.integers(1, 4)), dfx.fields.LinearDecay),
    "lotka_volterra": (lambda r: [1.5, -1.0, -3.0, 1.0], lambda r: 2, dfx.fields.LotkaVolterra),
    "lorenz": (lambda r: [10.0, 28.0, 8.0 / 3.0], lambda r: 3, dfx.fields.Lorenz),
    "vdp": (lambda r: [float(r.uniform(0.5, 3.0))], lambda r: 2, dfx.fields.VanDerPol),
    "forced_osc": (lambda r: [1.0, float(r.uniform(0.2, 1.0)), 2.0], lambda r: 2, dfx.fields.ForcedOscillator),
}
SOLVERS = {"tsit5": dfx.Tsit5, "dopri5": dfx.Dopri5, "dopri8": dfx.Dopri8, "heun": dfx.Heun, "bosh3": dfx.Bosh3,
           "midpoint": dfx.Midpoint, "ralston": dfx.Ralston, "euler": dfx.Euler, "shark": dfx.ShARK}
ADAPTIVE = ["tsit5", "dopri5", "dopri8", "heun", "bosh3", "half:heun", "half:euler", "half:midpoint", "half:ralston", "half:bosh3",
            "half:tsit5"]
FIXED = ["tsit5", "dopri5", "heun", "bosh3", "midpoint", "ralston", "euler", "half:euler"]


def make_solver(name):
    return dfx.HalfSolver(SOLVERS[name[5:]]()) if name.startswith("half:") else SOLVERS[name]()


def to_np(x):
    return x.detach().cpu().numpy() if isinstance(x, torch.Tensor) else np.asarray(x)


def relerr(a, b):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    if not np.array_equal(np.isfinite(a), np.isfinite(b)):
        return np.inf
    m = np.isfinite(b)
    if not m.any():
        return 0.0
    return float(np.max(np.abs(a[m] - b[m]) / (np.abs(b[m]) + 1e-3 * np.abs(b[m]).max() + 1e-300)))


def random_case(r):
    c = {}
    sde = r.random() < 0.25
    c["dtype"] = np.float32 if r.random() < 0.3 else np.float64
    n = int(r.choice([1, 2, 31, 32, 33, 100, 257]))
    c["n"] = n
    if sde:
        c["field"], c["params"], d = "ou", [float(r.uniform(0.5, 2)), float(r.uniform(-1, 1)), float(r.uniform(0.1, 0.8))], int(r.integers(1, 4))
        c["levy"] = "stla" if r.random() < 0.5 else "bi"
        fixed = r.random() < 0.6
        if fixed:
            c["solver"] = str(r.choice(["heun", "euler", "shark", "half:heun"] if c["levy"] == "stla" else ["heun", "euler", "half:heun"]))
        else:
            c["solver"] = str(r.choice(["heun", "half:heun", "half:shark"] if c["levy"] == "stla" else ["heun", "half:heun"]))
        c["bm_dim"] = 0 if (d == 1 and r.random() < 0.6) else d
        c["bm_tol"] = float(2.0 ** -int(r.integers(6, 11)))
        c["keys"] = r.integers(0, 2 ** 32, (n, 2), dtype=np.uint64).astype(np.uint32)
        c["t0"], c["t1"] = 0.0, 1.0
        if r.random() < 0.3:
            c["t0"], c["t1"] = 0.125, 0.875
    else:
        fname = str(r.choice(list(ODE_FIELDS)))
        pf, df, _ = ODE_FIELDS[fname]
        c["field"], c["params"], d = fname, pf(r), df(r)
        fixed = r.random() < 0.3
        c["solver"] = str(r.choice(FIXED if fixed else ADAPTIVE))
        c["t0"], c["t1"] = 0.0, float(r.uniform(0.5, 4.0))
        if r.random() < 0.25:  # backwards in time
            c["t0"], c["t1"] = c["t1"], c["t0"]
    c["d"] = d
    lo, hi = (0.5, 2.0)
    c["y0"] = r.uniform(lo, hi, (n, d)).astype(c["dtype"])
    span = abs(c["t1"] - c["t0"])
    sign = 1.0 if c["t1"] > c["t0"] else -1.0
    if fixed:
        c["controller"] = "constant"
        c["dt0"] = sign * span / float(r.choice([7, 16, 50, 64.5]))
    else:
        c["controller"] = "pid"
        f32 = c["dtype"] == np.float32
        c["rtol"] = float(10.0 ** r.uniform(-5 if f32 else -9, -3))
        c["atol"] = float(10.0 ** r.uniform(-6 if f32 else -10, -4))
        c["dt0"] = None if r.random() < 0.5 else sign * span * float(r.choice([0.01, 0.1, 1.0]))
        if c["dtype"] == np.float32 and c["solver"] == "dopri8" and c["dt0"] is None:
            # the same in fp32 for the 8th-order pair: the error estimate of a 0.01 first step is below eps
            c["dt0"] = sign * span * 0.1
        if c["solver"] in ("half:tsit5", "half:bosh3"):
            # step doubling around a 3rd / 5th order method: with the default first step of 0.01 the two results y1 and y1_alt
            # agree to the last place or differ by one ulp - |y1 - y1_alt| in {0, ulp} is pure rounding noise, and the reference's
            # factor is factormax for 0 (1 / 0 = inf, pid.py:498) but ~4 for one ulp.  Start these with a step that has an error
            c["dt0"] = sign * span * 0.25
        if r.random() < 0.4:
            c["pcoeff"], c["icoeff"], c["dcoeff"] = float(r.choice([0.0, 0.1, 0.4])), float(r.choice([1.0, 0.3])), float(r.choice([0.0, 0.05]))
        if r.random() < 0.15:
            c["dtmax"] = span / 8
        if r.random() < 0.1:
            c["dtmin"] = span / 64
            c["force_dtmin"] = bool(r.random() < 0.5)
    mode = int(r.integers(0, 6))
    c["save_t0"] = bool(r.random() < 0.3)
    c["save_t1"] = True
    if mode == 1:
        k = int(r.integers(1, 9))
        ts = np.sort(r.uniform(min(c["t0"], c["t1"]), max(c["t0"], c["t1"]), k))
        if r.random() < 0.3:
            ts[0], ts[-1] = min(c["t0"], c["t1"]), max(c["t0"], c["t1"])
        c["save_ts"] = (ts if sign > 0 else ts[::-1]).copy()
        c["save_t1"] = bool(r.random() < 0.5)
    elif mode == 2:
        c["save_steps"] = int(r.choice([1, 1, 2, 3]))
        c["save_t1"] = bool(r.random() < 0.5)
    elif mode == 3 and not sde and not c["solver"].startswith("half:"):
        c["save_dense"] = True
    c["max_steps"] = int(r.choice([4096, 4096, 256, 64, 12]))
    if c.get("save_steps") or c.get("save_dense"):
        c["max_steps"] = min(c["max_steps"], 256)
    if (not sde) and r.random() < 0.2:  # a per-trajectory end of the integration interval (vmapped t1)
        c["t1_per_traj"] = (c["t0"] + (c["t1"] - c["t0"]) * r.uniform(0.3, 1.0, n)).astype(c["dtype"])
        c.pop("save_ts", None)
        if not (c.get("save_steps") or c.get("save_dense")):
            c["save_t1"] = True
    if not (c["save_t0"] or c["save_t1"] or c.get("save_ts") is not None or c.get("save_steps") or c.get("save_dense")):
        c["save_t1"] = True
    c["host"] = bool(r.random() < 0.25)
    return c


def run_gpu(c, dev):
    field = dfx.fields.OrnsteinUhlenbeck(*c["params"]) if c["field"] == "ou" else ODE_FIELDS[c["field"]][2](*c["params"])
    y0 = c["y0"] if c["host"] else torch.tensor(c["y0"], device=dev)
    if c["controller"] == "constant":
        ctrl = dfx.ConstantStepSize()
    else:
        ctrl = dfx.PIDController(rtol=c["rtol"], atol=c["atol"], pcoeff=c.get("pcoeff", 0.0), icoeff=c.get("icoeff", 1.0),
                                 dcoeff=c.get("dcoeff", 0.0), dtmin=c.get("dtmin"), dtmax=c.get("dtmax"),
                                 force_dtmin=c.get("force_dtmin", True))
    saveat = dfx.SaveAt(t0=c["save_t0"], t1=c["save_t1"], ts=c.get("save_ts"), steps=c.get("save_steps", 0),
                        dense=c.get("save_dense", False))
    if c["field"] == "ou":
        keys = c["keys"] if c["host"] else torch.tensor(c["keys"].view(np.int32), device=dev)
        lv = dfx.BrownianIncrement if c["levy"] == "bi" else dfx.SpaceTimeLevyArea
        bm = dfx.VirtualBrownianTree(0.0, 1.0, c["bm_tol"], (c["bm_dim"],) if c["bm_dim"] else (), keys, lv)
        terms = dfx.MultiTerm(dfx.ODETerm(field.drift), dfx.ControlTerm(field.diffusion, bm))
    else:
        terms = dfx.ODETerm(field)
    t1 = c["t1"]
    if "t1_per_traj" in c:
        t1 = c["t1_per_traj"] if c["host"] else torch.tensor(c["t1_per_traj"], device=dev)
    return dfx.diffeqsolve(terms, make_solver(c["solver"]), c["t0"], t1, c["dt0"], y0, saveat=saveat,
                           stepsize_controller=ctrl, max_steps=c["max_steps"], throw=False)


def run_oracle(c, **over):
    kw = dict(solver=c["solver"], params=c["params"], dtype=c["dtype"], controller=c["controller"],
              save_t0=c["save_t0"], save_t1=c["save_t1"], save_ts=c.get("save_ts"), save_steps=c.get("save_steps", 0),
              save_dense=c.get("save_dense", False), max_steps=c["max_steps"])
    if c["controller"] == "pid":
        kw.update(rtol=c["rtol"], atol=c["atol"], pcoeff=c.get("pcoeff", 0.0), icoeff=c.get("icoeff", 1.0), dcoeff=c.get("dcoeff", 0.0),
                  dtmin=c.get("dtmin"), dtmax=c.get("dtmax"), force_dtmin=c.get("force_dtmin", True))
    if c["field"] == "ou":
        kw.update(levy_area=c["levy"], keys=c["keys"], bm_t0=0.0, bm_t1=1.0, bm_tol=c["bm_tol"], bm_dim=c["bm_dim"])
    if "t1_per_traj" in c:
        kw["t1_per_traj"] = c["t1_per_traj"]
    kw.update(over)
    return oracle.solve(c["field"], c["y0"], c["t0"], c["t1"], c["dt0"], **kw)


def sensitivity2(c, o):
    """Second tier, adaptive solves only.  The embedded error estimate is a difference of nearly equal numbers (for HalfSolver
    literally |y1 - y1_alt|), so its RELATIVE rounding noise is eps |y| / |err| - 1e-6 and more in fp64 when a step is far more
    accurate than asked for - and it differs between the GPU's contracted FMAs and the oracle.  A 1-ulp probe of y0 moves
    y1 and y1_alt together and does not see it; scaling the tolerances by (1 +- 1e-6) (fp64) / (1 +- 1e-3) (fp32) does: it is the
    same perturbation of err / tolerance."""
    if c["controller"] != "pid":
        return 1.0, 0.0, 0.0
    h = 1e-3 if c["dtype"] == np.float32 else 1e-6
    frac, sy, st = 1.0, 0.0, 0.0
    for sgn in (1.0, -1.0):
        f = run_oracle(c, rtol=c["rtol"] * (1 + sgn * h), atol=c["atol"] * (1 + sgn * h))
        same = np.all(f["stats"] == o["stats"], axis=1) & (f["result"] == o["result"])
        if not same.any():
            return 0.0, np.inf, np.inf
        frac = min(frac, float(same.mean()))
        sy, st = max(sy, relerr(f["ys"][same], o["ys"][same])), max(st, relerr(f["ts"][same], o["ts"][same]))
    return frac, sy, st


def sensitivity(c, o):
    """The oracle's own rounding sensitivity on this case: strict IEEE sequencing against the FMA-contracted build (the two
    differ the way the CUDA compiler's contraction differs from the oracle).  Returns (fraction of trajectories whose step
    statistics agree between the two builds, rel. difference of ys, of ts on those)."""
    with oracle.rounding("fma"):
        probes = [run_oracle(c)]
    if c["controller"] == "pid":
        # ... and to a last-place change in the controller (pow() differs in the last place between libm and the GPU): `safety` one ulp up / down
        probes += [run_oracle(c, safety=float(np.nextafter(0.9, 1.0))), run_oracle(c, safety=float(np.nextafter(0.9, 0.0)))]
    # ... and to one ulp in the initial condition (moves every rounding of the solve, the error estimate's cancellation included)
    for up in (np.inf, -np.inf):
        probes.append(run_oracle({**c, "y0": np.nextafter(c["y0"], np.asarray(up, c["dtype"]))}))
    frac, sy, st = 1.0, 0.0, 0.0
    for f in probes:
        same = np.all(f["stats"] == o["stats"], axis=1) & (f["result"] == o["result"])
        if not same.any():
            return 0.0, np.inf, np.inf
        frac = min(frac, float(same.mean()))
        sy, st = max(sy, relerr(f["ys"][same], o["ys"][same])), max(st, relerr(f["ts"][same], o["ts"][same]))
    return frac, sy, st


TIER2 = []


def check(c, sol, o):
    msgs = check_raw(c, sol, o)
    if not msgs:
        return msgs
    frac, sy, st_ = sensitivity(c, o)
    st = np.stack([to_np(sol.stats[k]) for k in ("num_steps", "num_accepted_steps", "num_rejected_steps")], 1)
    same = np.all(st == o["stats"], axis=1) & (to_np(sol.result) == o["result"])
    ys, ts = to_np(sol.ys), to_np(sol.ts)
    e_y = relerr(ys[same], o["ys"].reshape(ys.shape)[same]) if same.any() else 0.0
    e_t = relerr(ts[same], o["ts"][same]) if same.any() else 0.0
    # explained by rounding sensitivity: the two oracle builds disagree about as much as the GPU and the oracle do
    if e_y <= 64 * sy + 1e-300 and e_t <= 64 * st_ + (1e-6 if c["dtype"] == np.float32 else 1e-13) and same.mean() >= min(0.9, frac) - 0.25:
        return []
    frac2, sy2, st2 = sensitivity2(c, o)
    if e_y <= 8 * sy2 + 1e-300 and e_t <= 8 * st2 + (1e-6 if c["dtype"] == np.float32 else 1e-13) and same.mean() >= min(0.9, frac2) - 0.25:
        TIER2.append(1)
        return []
    return msgs + [f"(oracle vs its own fma / safety+-1ulp / y0+-1ulp probes: stats agree {frac:.3f}, ys {sy:.2e}, ts {st_:.2e}; gpu vs oracle: stats agree {same.mean():.3f}, ys {e_y:.2e}, ts {e_t:.2e})"]


def check_raw(c, sol, o):
    f32 = c["dtype"] == np.float32
    st = np.stack([to_np(sol.stats[k]) for k in ("num_steps", "num_accepted_steps", "num_rejected_steps")], 1)
    res = to_np(sol.result)
    same = np.all(st == o["stats"], axis=1) & (res == o["result"])
    # adaptive solves: a last-ulp difference can flip one accept / reject decision; fixed-step ones must agree everywhere
    need = 1.0 if c["controller"] == "constant" else (0.5 if f32 else 0.9)
    msgs = []
    if same.mean() < need and not (c["n"] <= 2 and not f32 and c["controller"] == "pid" and same.mean() >= 0.5):
        msgs.append(f"statistics/result agree on {same.mean():.3f} of the trajectories (need {need})")
    if c["controller"] == "pid" and np.abs(st[:, 1] - o["stats"][:, 1]).max() > (3 if f32 else 2):
        msgs.append(f"accepted steps differ by up to {np.abs(st[:, 1] - o['stats'][:, 1]).max()}")
    tol = 2e-4 if f32 else 1e-9
    if c["field"] == "ou" and c["controller"] == "pid":
        tol = 1e-3 if f32 else 1e-5   # adaptive stepping on a Brownian path: chaotic in the step times even at equal counts
    if same.any():
        ys, ts = to_np(sol.ys), to_np(sol.ts)
        oys = o["ys"].reshape(ys.shape)
        e_y, e_t = relerr(ys[same], oys[same]), relerr(ts[same], o["ts"][same])
        if not e_y < tol:
            msgs.append(f"ys rel err {e_y:.3e} >= {tol}")
        if not e_t < (1e-5 if f32 else 1e-12):
            msgs.append(f"ts rel err {e_t:.3e}")
    return msgs


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cases", type=int, default=300)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--verbose", action="store_true")
    ap.add_argument("--prebuild", action="store_true", help="no GPU needed: compile (in parallel) the kernels that the cases will "
                    "instantiate on first use, so that a following GPU run finds them under diffrax_b200/lib/user/")
    a = ap.parse_args()
    if a.prebuild:
        import concurrent.futures as cf
        r = np.random.default_rng(a.seed)
        cases = [random_case(r) for _ in range(a.cases)]

        def one(c):
            try:
                run_gpu({**c, "host": True}, None)
            except Exception as e:  # noqa: BLE001  (without a GPU every case ends in "CUDA device not available" - after the build)
                return str(e)[:80]
            return "ok"
        with cf.ThreadPoolExecutor(os.cpu_count() or 4) as ex:
            out = list(ex.map(one, cases))
        print("prebuild:", {k: out.count(k) for k in sorted(set(out))})
        return
    dev = torch.device("cuda:0")
    r = np.random.default_rng(a.seed)
    bad = skipped = 0
    refusals = {}
    for i in range(a.cases):
        c = random_case(r)
        tag = {k: (v if not isinstance(v, np.ndarray) else f"<{v.shape}>") for k, v in c.items() if k not in ("y0", "keys")}
        tag["dtype"] = np.dtype(c["dtype"]).name
        try:
            sol = run_gpu(c, dev)
            torch.cuda.synchronize()
        except (ValueError, RuntimeError, NotImplementedError) as e:   # a combination the facade refuses (as the reference would) or has no kernel for
            skipped += 1
            key = f"{c['field']}/d{c['d']}/{c['solver']}/{np.dtype(c['dtype']).name}/{c.get('levy')}: {str(e)[:90]}"
            refusals[key] = refusals.get(key, 0) + 1
            if a.verbose:
                print(f"[{i}] refused: {str(e)[:100]}  {tag}")
            continue
        try:
            o = run_oracle(c)
            msgs = check(c, sol, o)
        except Exception:
            msgs = ["exception: " + traceback.format_exc(limit=3)]
        if msgs:
            bad += 1
            print(f"[{i}] FAIL {msgs}  {tag}", flush=True)
        elif a.verbose:
            print(f"[{i}] ok {tag}")
    for k, v in sorted(refusals.items()):
        print(f"refused x{v}: {k}")
    print(f"fuzz_parity: {a.cases} cases, {skipped} refused by the facade, {bad} failures (seed {a.seed}); "
          f"{len(TIER2)} cases needed the tolerance-scaling probe (error-estimate cancellation noise)")
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
