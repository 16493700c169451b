"""Ad-hoc GPU shake-out: CUDA path vs oracle on a handful of cases (not a test; see tests/)."""
import sys, time, json
sys.path.insert(0, ".")
import numpy as np, torch
import diffrax_b200 as dfx
import oracle

dev = torch.device("cuda:0")
print(torch.cuda.get_device_name(0))

def rel(a, b):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    m = np.isfinite(a) & np.isfinite(b)
    same_inf = np.array_equal(np.isfinite(a), np.isfinite(b))
    if m.sum() == 0: return 0.0, same_inf
    scale = np.abs(b[m]).max()
    return float(np.max(np.abs(a[m] - b[m]) / (np.abs(b[m]) + 1e-3 * scale))), same_inf

rng = np.random.default_rng(1)
N = 4096
# --- C2: Lorenz Dopri5 ---
y0 = np.stack([rng.uniform(-15, 15, N), rng.uniform(-20, 20, N), rng.uniform(5, 45, N)], 1)
o = oracle.solve("lorenz", y0, 0.0, 2.0, None, solver="dopri5", params=[10., 28., 8/3], rtol=1e-8, atol=1e-8)
s = dfx.diffeqsolve(dfx.ODETerm(dfx.fields.Lorenz()), dfx.Dopri5(), 0.0, 2.0, None, torch.tensor(y0, device=dev),
                    stepsize_controller=dfx.PIDController(1e-8, 1e-8))
torch.cuda.synchronize()
ys = s.ys.cpu().numpy(); st = torch.stack([s.stats[k] for k in ("num_steps","num_accepted_steps","num_rejected_steps")],1).cpu().numpy()
print("C2 lorenz dopri5: max rel", rel(ys, o["ys"]), "stats equal", np.array_equal(st, o["stats"]), "max |dsteps|", np.abs(st-o["stats"]).max(), "mean steps", st[:,0].mean())
# host path
s2 = dfx.diffeqsolve(dfx.ODETerm(dfx.fields.Lorenz()), dfx.Dopri5(), 0.0, 2.0, None, y0, stepsize_controller=dfx.PIDController(1e-8, 1e-8))
print("  host path identical:", np.array_equal(s2.ys, ys))

# --- C1: LV Tsit5 ts ---
N1 = 1024
y0 = rng.uniform(0.5, 2, (N1, 2)); ts = np.linspace(0, 10, 100)
o = oracle.solve("lotka_volterra", y0, 0.0, 10.0, None, solver="tsit5", params=[1.5,-1,-3,1], rtol=1e-6, atol=1e-6, save_t1=False, save_ts=ts)
s = dfx.diffeqsolve(dfx.ODETerm(dfx.fields.LotkaVolterra()), dfx.Tsit5(), 0.0, 10.0, None, torch.tensor(y0, device=dev),
                    saveat=dfx.SaveAt(ts=ts), stepsize_controller=dfx.PIDController(1e-6, 1e-6))
print("C1 LV tsit5 ts: ys", rel(s.ys.cpu().numpy(), o["ys"]), "ts", rel(s.ts.cpu().numpy(), o["ts"]), "steps equal", np.array_equal(s.stats["num_steps"].cpu().numpy(), o["stats"][:,0]))

# --- every solver, every save mode on a small problem ---
for name, cls in [("tsit5", dfx.Tsit5), ("dopri5", dfx.Dopri5), ("dopri8", dfx.Dopri8), ("heun", dfx.Heun), ("bosh3", dfx.Bosh3), ("midpoint", dfx.Midpoint), ("ralston", dfx.Ralston)]:
    y0 = rng.uniform(-2, 2, (64, 2)); tsv = np.linspace(0.0, 3.0, 7)
    kw = dict(rtol=1e-5, atol=1e-7)
    o = oracle.solve("forced_osc", y0, 0.0, 3.0, None, solver=name, params=[1.0, 0.7, 2.0], save_t0=True, save_t1=True, save_ts=tsv, save_steps=2, save_dense=True, max_steps=512, **kw)
    s = dfx.diffeqsolve(dfx.ODETerm(dfx.fields.ForcedOscillator(1.0, 0.7, 2.0)), cls(), 0.0, 3.0, None, torch.tensor(y0, device=dev),
                        saveat=dfx.SaveAt(t0=True, t1=True, ts=tsv, steps=2, dense=True), stepsize_controller=dfx.PIDController(**kw), max_steps=512, throw=False)
    di = s.interpolation
    q = np.linspace(0, 3, 11)
    ev = di.evaluate(q).cpu().numpy()
    oev = oracle.dense_evaluate(name, o["dense"], np.tile(q, (64, 1)))
    print(f"{name:9s} ys {rel(s.ys.cpu().numpy(), o['ys'])} ts {rel(s.ts.cpu().numpy(), o['ts'])} steps_eq {np.array_equal(s.stats['num_steps'].cpu().numpy(), o['stats'][:,0])} "
          f"res_eq {np.array_equal(s.result.cpu().numpy(), o['result'])} dense_ts {rel(di.ts.cpu().numpy(), o['dense']['ts'])} dense_k {rel(di.infos['k'].cpu().numpy(), o['dense']['k'])} eval {rel(ev, oev)}")

# --- PRNG ---
from diffrax_b200 import _lib
L = _lib.lib()
keys = torch.tensor(np.array([[0,0],[0xffffffff,0xffffffff],[0x13198a2e,0x03707344]], np.uint32).view(np.int32), device=dev)
ctrs = torch.tensor(np.array([[0,0],[0xffffffff,0xffffffff],[0x243f6a88,0x85a308d3]], np.uint32).view(np.int32), device=dev)
out = torch.empty_like(keys)
L.dfx_threefry2x32(3, keys.data_ptr(), ctrs.data_ptr(), out.data_ptr(), None); torch.cuda.synchronize()
print("threefry KAT:", [hex(v) for v in out.cpu().numpy().view(np.uint32).ravel()])
kk = dfx.random.split(dfx.random.key(42), 100000)
kd = torch.tensor(kk.view(np.int32), device=dev)
for dt_, odt, did in [(torch.float64, np.float64, 0), (torch.float32, np.float32, 1)]:
    for part in (1, 0):
        z = torch.empty(kk.shape[0], dtype=dt_, device=dev)
        L.dfx_random_normal(did, kk.shape[0], kd.data_ptr(), part, z.data_ptr(), None); torch.cuda.synchronize()
        zo = np.array([oracle.normal(k, odt, bool(part)) for k in kk[:20000]])
        zz = z.cpu().numpy()[:20000]
        ulp = np.abs(zz - zo) / np.spacing(np.abs(zo).astype(odt))
        print(f"normal {odt.__name__} part={part}: bit-equal {np.mean(zz == zo):.4f} max ulp {ulp.max():.1f} mean {zz.mean():.4f} var {zz.var():.4f}")
# --- VBT increments ---
for lv, cls in (("bi", dfx.BrownianIncrement), ("stla", dfx.SpaceTimeLevyArea)):
    for dt_, odt in ((torch.float64, np.float64), (torch.float32, np.float32)):
        bm = dfx.VirtualBrownianTree(0.0, 1.0, 2**-8, (), kd[:20000], cls)
        ta = torch.full((20000,), 0.3, dtype=dt_, device=dev); tb = torch.full((20000,), 0.7, dtype=dt_, device=dev)
        W, H = bm.evaluate(ta, tb, use_levy=True)
        Wo, Ho = oracle.vbt_evaluate(kk[:20000], 0.3, 0.7, tol=2**-8, levy_area=lv, dtype=odt)
        print(f"vbt {lv} {odt.__name__}: W bit-equal {np.mean(W.cpu().numpy()==Wo):.4f} maxabs {np.abs(W.cpu().numpy()-Wo).max():.2e}  H bit-equal {np.mean(H.cpu().numpy()==Ho):.4f} maxabs {np.abs(H.cpu().numpy()-Ho).max():.2e}")
# --- C5: OU ---
N5 = 20000
for sname, cls, lv, lcls in (("heun", dfx.Heun, "bi", dfx.BrownianIncrement), ("shark", dfx.ShARK, "stla", dfx.SpaceTimeLevyArea), ("euler", dfx.Euler, "bi", dfx.BrownianIncrement)):
    for dt_, odt in ((torch.float64, np.float64), (torch.float32, np.float32)):
        ou = dfx.fields.OrnsteinUhlenbeck(1.0, 0.0, 0.5)
        bm = dfx.VirtualBrownianTree(0.0, 1.0, 2**-8, (), kd[:N5], lcls)
        s = dfx.diffeqsolve(dfx.MultiTerm(dfx.ODETerm(ou.drift), dfx.ControlTerm(ou.diffusion, bm)), cls(), 0.0, 1.0, 2**-6, torch.ones(N5, 1, dtype=dt_, device=dev))
        o = oracle.solve("ou", np.ones((N5, 1)), 0.0, 1.0, 2**-6, solver=sname, params=[1.0, 0.0, 0.5], dtype=odt, controller="constant", levy_area=lv, keys=kk[:N5], bm_tol=2**-8)
        a = s.ys.cpu().numpy(); b = o["ys"]
        print(f"OU {sname} {odt.__name__}: max abs diff {np.abs(a-b).max():.3e} bit-equal {np.mean(a==b):.4f} steps {s.stats['num_steps'][0].item()} mean {a.mean():.4f} var {a.var():.4f}")
# --- timing C2 at scale ---
N = 1 << 20
y0 = torch.tensor(np.stack([rng.uniform(-15, 15, N), rng.uniform(-20, 20, N), rng.uniform(5, 45, N)], 1), device=dev)
term = dfx.ODETerm(dfx.fields.Lorenz()); ctl = dfx.PIDController(1e-8, 1e-8)
for it in range(3):
    e0 = torch.cuda.Event(True); e1 = torch.cuda.Event(True)
    e0.record(); s = dfx.diffeqsolve(term, dfx.Dopri5(), 0.0, 2.0, None, y0, stepsize_controller=ctl); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1); acc = int(s.stats["num_accepted_steps"].sum()); att = int(s.stats["num_steps"].sum())
    print(f"C2 N=2^20: {ms:.2f} ms  accepted {acc} attempted {att}  -> {acc/ms*1e3:.3e} acc steps/s, {att*316/ms*1e3/1e12:.2f} TFLOP/s algorithmic")
print("fp64 fma peak TF/s", L.dfx_measure_fma_peak(0, 0), "fp32", L.dfx_measure_fma_peak(1, 0), "int Tops", L.dfx_measure_int_peak(0))
