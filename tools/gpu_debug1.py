import sys; sys.path.insert(0, "."); sys.path.insert(0, "tests"); sys.path.insert(0, "tests/golden")
import numpy as np, torch
import diffrax_b200 as dfx, oracle
from test_gpu_parity import run_case, stats_np, to_np, relerr
dev = torch.device("cuda:0")
for solver, dtype in (("tsit5", np.float64), ("dopri5", np.float64), ("dopri8", np.float64), ("tsit5", np.float32), ("heun", np.float32)):
    rng = np.random.default_rng(11); n = 200
    tight = solver in ("tsit5", "dopri5", "dopri8"); f32 = dtype == np.float32
    kw = dict(field="forced_osc", params=[1.0, 0.7, 2.0], solver=solver, dtype=dtype, y0=rng.uniform(-2, 2, (n, 2)).astype(dtype),
              t0=0.0, t1=3.0, dt0=None, rtol=(1e-4 if f32 else (1e-9 if tight else 1e-5)), atol=(1e-6 if f32 else (1e-11 if tight else 1e-7)),
              save_t0=True, save_t1=True, save_ts=np.linspace(0.0, 3.0, 13), save_steps=2, save_dense=True, max_steps=2048)
    o = oracle.solve(kw["field"], kw["y0"], 0.0, 3.0, None, **{k: v for k, v in kw.items() if k not in ("field", "y0", "t0", "t1", "dt0")})
    sol = run_case(kw, dev)
    st = stats_np(sol)
    diff = np.where(np.any(st != o["stats"], axis=1))[0]
    print(f"== {solver} {dtype.__name__}: {len(diff)} of {n} trajectories differ in stats; max d_acc {np.abs(st[:,1]-o['stats'][:,1]).max()}")
    same = np.all(st == o["stats"], axis=1)
    dts_g = to_np(sol.interpolation.ts); dts_o = o["dense"]["ts"]
    print("   same-traj: dense_ts relerr", relerr(dts_g[same], dts_o[same]), " ys(all) relerr", relerr(to_np(sol.ys)[same], o["ys"][same]))
    ots = o["ts"][same]; fixed = np.isin(ots, np.linspace(0.0, 3.0, 13).astype(dtype))
    print("   fixed-time outputs relerr", relerr(np.where(fixed[..., None], to_np(sol.ys)[same], 0), np.where(fixed[..., None], o["ys"][same], 0)))
    for i in diff[:3]:
        a, b = dts_g[i], dts_o[i]
        m = min(np.isfinite(a).sum(), np.isfinite(b).sum())
        rel = np.abs(a[:m] - b[:m]) / np.maximum(np.abs(b[:m]), 1e-30)
        first = np.argmax(rel > (1e-3 if f32 else 1e-6)) if np.any(rel > (1e-3 if f32 else 1e-6)) else -1
        print(f"   traj {i}: gpu stats {st[i]} oracle {o['stats'][i]} first divergent knot {first} rel there {rel[first] if first>=0 else 0:.3e}; rel just before {rel[max(first-1,0)]:.3e}")
        if first > 1:
            print("      gpu knots", a[first-2:first+2], "\n      orc knots", b[first-2:first+2])
