"""Warp-per-trajectory kernel (csrc/wide_kernel.cuh) on Lorenz-96: D components, N trajectories, Dopri5 / PID(1e-8) / fp64.
    python tools/wide_bench.py [D] [N]
Flop count per attempted step (SURVEY 8d formula with the field at 4 flop per component): D * ((s-1) * 4 + s(s-1) + 2(s-1) + 2s) + 8D."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import diffrax_b200 as dfx  # noqa: E402

D = int(sys.argv[1]) if len(sys.argv) > 1 else 128
N = int(sys.argv[2]) if len(sys.argv) > 2 else 1 << 15
L96 = "fi = (y[(i + 1) % D] - y[(i + D - 2) % D]) * y[(i + D - 1) % D] - y[i] + p[0];"
MB = os.environ.get("WIDE_MIN_BLOCKS")
f = dfx.fields.CudaField(D, L96, params=[8.0], wide=True, defines={"DFX_WIDE_MIN_BLOCKS": int(MB)} if MB else None)
rng = np.random.default_rng(0)
y0 = torch.tensor(8.0 + rng.normal(0, 0.5, (N, D)), device="cuda")
plan = dfx.prepare(dfx.ODETerm(f), dfx.Dopri5(), 0.0, 1.0, None, y0, stepsize_controller=dfx.PIDController(rtol=1e-8, atol=1e-8))
for _ in range(3):
    sol = plan(throw=False)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
K = 10
e0.record()
for _ in range(K):
    sol = plan(throw=False)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / K
att, acc = int(sol.stats["num_steps"].sum()), int(sol.stats["num_accepted_steps"].sum())
s = 7
flop = D * ((s - 1) * 4 + s * (s - 1) + 2 * (s - 1) + 2 * s) + 8 * D
print(f"[min_blocks {MB}] Lorenz-96 D={D} N={N} Dopri5 fp64: {ms:.3f} ms per solve, {acc / ms * 1e3:.3e} accepted steps/s, "
      f"{att * flop / ms * 1e3 / 1e12:.2f} TFLOP/s algorithmic ({flop} flop per attempted step), failed {int((sol.result != 0).sum())}")
