"""Sweep DFX_HOST_CHUNKS for the pipelined host entry point on C2 (one process, median of 15 solves per setting)."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import diffrax_b200 as dfx  # noqa: E402

n = 1 << 20
rng = np.random.default_rng(1)
y0 = torch.tensor(np.stack([rng.uniform(-15, 15, n), rng.uniform(-20, 20, n), rng.uniform(5, 45, n)], 1)).pin_memory()
term, ctrl = dfx.ODETerm(dfx.fields.Lorenz(10.0, 28.0, 8.0 / 3.0)), dfx.PIDController(rtol=1e-8, atol=1e-8)


def solve():
    s = dfx.diffeqsolve(term, dfx.Dopri5(), 0.0, 2.0, None, y0, stepsize_controller=ctrl, throw=False, device=0)
    return int(s.result[0])


for chunks in sys.argv[1:] or ["16", "32", "64", "8"]:
    os.environ["DFX_HOST_CHUNKS"] = chunks
    for _ in range(3):
        solve()
    ts = []
    for _ in range(15):
        t = time.perf_counter()
        solve()
        ts.append(time.perf_counter() - t)
    print(f"chunks {chunks}: median {np.median(ts) * 1e3:.3f} ms, min {min(ts) * 1e3:.3f} ms", flush=True)
