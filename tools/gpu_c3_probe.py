import sys; sys.path.insert(0, ".")
import numpy as np, torch, diffrax_b200 as dfx
dev = torch.device("cuda:0")
rng = np.random.default_rng(2); n = 1 << 18
y0 = torch.tensor(np.array([0.994, 0.0, 0.0, -2.00158510637908252]) + 1e-4 * rng.standard_normal((n, 4)), device=dev)
term, ctrl = dfx.ODETerm(dfx.fields.CR3BP()), dfx.PIDController(1e-12, 1e-12)
for name, sa in (("t1", dfx.SaveAt(t1=True)), ("steps", dfx.SaveAt(steps=True)), ("dense", dfx.SaveAt(dense=True))):
    plan = dfx.prepare(term, dfx.Dopri8(), 0.0, 17.0652165601579625, None, y0, saveat=sa, stepsize_controller=ctrl, max_steps=768)
    for _ in range(2): plan(throw=False)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record(); s = plan(throw=False); e1.record(); torch.cuda.synchronize()
    print(name, f"{e0.elapsed_time(e1):.2f} ms", "attempted", int(s.stats["num_steps"].sum()), "failed", int((s.result != 0).sum()))
