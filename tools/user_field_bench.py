"""A user-written Lorenz field (fields.CudaField) against the built-in functor on the C2 workload: same bits, same speed?
    python tools/user_field_bench.py"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import diffrax_b200 as dfx  # noqa: E402

SRC = "f[0] = p[0] * (y[1] - y[0]); f[1] = y[0] * (p[1] - y[2]) - y[1]; f[2] = y[0] * y[1] - p[2] * y[2];"
rng = np.random.default_rng(1)
n = 1 << 20
y0 = torch.tensor(np.stack([rng.uniform(-15, 15, n), rng.uniform(-20, 20, n), rng.uniform(5, 45, n)], 1), device="cuda")
ctrl = dfx.PIDController(rtol=1e-8, atol=1e-8)
fields = {"built-in Lorenz": dfx.fields.Lorenz(), "CudaField (heuristic occupancy)": dfx.fields.CudaField(3, SRC, params=[10.0, 28.0, 8.0 / 3.0]),
          "CudaField(min_blocks_per_sm=6)": dfx.fields.CudaField(3, SRC, params=[10.0, 28.0, 8.0 / 3.0], min_blocks_per_sm=6)}
fields["CudaField(min_blocks_per_sm=6) + per-trajectory args [N, 3]"] = fields["CudaField(min_blocks_per_sm=6)"]
args = torch.tensor(np.tile([10.0, 28.0, 8.0 / 3.0], (n, 1)), device="cuda")
ref = None
for name, f in fields.items():
    plan = dfx.prepare(dfx.ODETerm(f), dfx.Dopri5(), 0.0, 2.0, None, y0, args if "args" in name else None, stepsize_controller=ctrl)
    for _ in range(3):
        sol = plan(throw=False)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        sol = plan(throw=False)
    e1.record(); torch.cuda.synchronize()
    ys = sol.ys.clone()
    ref = ys if ref is None else ref
    print(f"{name:62s} {e0.elapsed_time(e1) / 20:.3f} ms per solve of 2^20 trajectories; bit-identical to the built-in: {bool(torch.equal(ys, ref))}")
