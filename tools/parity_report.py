#!/usr/bin/env python
"""Prints, for every committed golden case, how far the CUDA path is from the oracle: accepted-step differences, fraction of
identical step statistics, max relative state error.  (GPU box; the numbers behind the tolerances in tests/test_gpu_parity.py.)"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import make_golden  # noqa: E402
import test_gpu_parity as T  # noqa: E402

dev = torch.device("cuda:0")
print(f"{'case':28s} {'max|d acc|':>10s} {'same stats':>10s} {'max rel err (same)':>20s} {'state-vector err':>18s}")
for name, kw in make_golden.CASES.items():
    sol = T.run_case(kw, dev)
    st, gst = T.stats_np(sol), T.GOLD[f"{name}/stats"]
    same = np.all(st == gst, axis=1)
    ys, gys = T.to_np(sol.ys), T.GOLD[f"{name}/ys"]
    if ys.ndim == 2:
        ys = ys[..., None]
    gys = gys.reshape(ys.shape)
    if kw.get("save_steps"):
        e1 = e2 = float("nan")
    else:
        e1 = T.relerr(ys[same], gys[same]) if same.any() else float("nan")
        e2 = T.relerr_state(ys[same], gys[same]) if same.any() else float("nan")
    print(f"{name:28s} {np.abs(st[:, 1] - gst[:, 1]).max():10d} {same.mean():10.3f} {e1:20.3e} {e2:18.3e}")
