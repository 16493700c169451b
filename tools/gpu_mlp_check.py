"""TC kernel vs CUDA-core functor vs oracle for the C4 neural-ODE field."""
import os, sys, time
sys.path.insert(0, ".")
import numpy as np, torch
import diffrax_b200 as dfx, oracle
dev = torch.device("cuda:0")
mlp = dfx.fields.MLP.init(3, d=4, width=128)
rng = np.random.default_rng(30)
n = int(os.environ.get("N", 4096))
y0 = rng.standard_normal((n, 4)).astype(np.float32)
term, ctrl = dfx.ODETerm(mlp), dfx.PIDController(rtol=1e-3, atol=1e-6)
def run(tag):
    y0d = torch.tensor(y0, device=dev)
    plan = dfx.prepare(term, dfx.Tsit5(), 0.0, 10.0, None, y0d, stepsize_controller=ctrl)
    s = plan(throw=False); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record(); s = plan(throw=False); e1.record(); torch.cuda.synchronize()
    print(f"{tag}: {e0.elapsed_time(e1):.3f} ms  steps {int(s.stats['num_steps'].sum())} failed {int((s.result != 0).sum())}")
    return s.ys.cpu().numpy().copy(), torch.stack([s.stats[k] for k in ("num_steps", "num_accepted_steps")], 1).cpu().numpy().copy()
mode = sys.argv[1] if len(sys.argv) > 1 else "tc"
if mode == "cc": os.environ["DFX_MLP_NO_TC"] = "1"
if mode == "tc_exact": os.environ["DFX_MLP_EXACT_ACT"] = "1"
ys, st = run(mode)
o = oracle.solve("mlp", y0, 0.0, 10.0, None, solver="tsit5", params=mlp.oracle_params(), dtype=np.float32, rtol=1e-3, atol=1e-6)
nb = np.linalg.norm(o["ys"][:, 0], axis=-1)
err = np.linalg.norm(ys[:, 0] - o["ys"][:, 0], axis=-1) / (nb + 1e-3 * nb.max())
same = np.all(st == o["stats"][:, :2], axis=1)
print(f"{mode}: same stats {same.mean():.4f}  max rel state err (same) {err[same].max() if same.any() else -1:.3e}  (all) {err.max():.3e}  median {np.median(err):.3e}")
print("sample", ys[0, 0], o["ys"][0, 0])
