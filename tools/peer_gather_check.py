#!/usr/bin/env python
"""torchrun check of the fused peer-memory gather (needs >= 2 GPUs on one node):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 tools/peer_gather_check.py

Every rank solves its block of a global Lorenz batch twice - once with gather="peer" (finals stored into every rank's buffer
by the solve kernel over NVLink, then a 32-byte all_gather) and once with gather="nccl" (packed record all_gather) - on device
AND host inputs, and checks that both give every rank the same global finals / statistics as a single-GPU solve of the batch."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import diffrax_b200 as dfx  # noqa: E402
from diffrax_b200 import _dist  # noqa: E402

rank, local, world = _dist.init_from_env()
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
n = 50001                       # uneven blocks
rng = np.random.default_rng(3)
y0 = np.stack([rng.uniform(-15, 15, n), rng.uniform(-20, 20, n), rng.uniform(5, 45, n)], 1)
term, ctrl = dfx.ODETerm(dfx.fields.Lorenz()), dfx.PIDController(1e-6, 1e-6)
ref = dfx.diffeqsolve(term, dfx.Dopri5(), 0.0, 1.0, None, torch.tensor(y0, device=dev), stepsize_controller=ctrl)
for host in (False, True):
    inp = torch.tensor(y0).pin_memory() if host else torch.tensor(y0, device=dev)
    for mode in ("peer", "nccl"):
        plan = dfx.prepare_sharded(term, dfx.Dopri5(), 0.0, 1.0, None, inp, stepsize_controller=ctrl, gather=mode, device=dev)
        for it in range(3):     # the two peer buffers alternate
            out = plan()
            torch.cuda.synchronize()
            assert torch.equal(out.y_final, ref.ys[:, 0]), (rank, mode, host, it)
            assert bool((out.t_final == 1.0).all())
            assert int(out.stats["num_steps"]) == int(ref.stats["num_steps"].sum()), (rank, mode)
            assert int(out.stats["num_accepted_steps"]) == int(ref.stats["num_accepted_steps"].sum())
            assert int(out.stats["num_failed"]) == 0 and int(out.stats["max_steps_per_trajectory"]) == int(ref.stats["num_steps"].max())
        plan.close()
# fp32 SDE with per-trajectory keys through the peer gather
keys = dfx.random.split(dfx.random.key(9), n)
kd = torch.tensor(keys.view(np.int32), device=dev)
ou = dfx.fields.OrnsteinUhlenbeck(1.0, 0.0, 0.5)
mk = lambda k: dfx.MultiTerm(dfx.ODETerm(ou.drift), dfx.ControlTerm(ou.diffusion, dfx.VirtualBrownianTree(0.0, 1.0, 2.0 ** -7, (), k)))  # noqa: E731
y1 = torch.ones(n, 1, dtype=torch.float32, device=dev)
r2 = dfx.diffeqsolve(mk(kd), dfx.Heun(), 0.0, 1.0, 2.0 ** -5, y1)
p2 = dfx.prepare_sharded(mk(kd), dfx.Heun(), 0.0, 1.0, 2.0 ** -5, y1, gather="peer")
o2 = p2()
torch.cuda.synchronize()
assert torch.equal(o2.y_final, r2.ys[:, 0]) and int(o2.stats["num_steps"]) == 32 * n
p2.close()
# a wide state (one trajectory per warp, csrc/wide_kernel.cuh): the same fused gather, every rank compiles / loads the plugin
D, nw = 40, 3001
l96 = dfx.fields.CudaField(D, "fi = (y[(i + 1) % D] - y[(i + D - 2) % D]) * y[(i + D - 1) % D] - y[i] + p[0];", params=[8.0], wide=True)
yw = torch.tensor(8.0 + np.random.default_rng(5).normal(0, 0.5, (nw, D)), device=dev)
cw = dfx.PIDController(1e-7, 1e-7)
r3 = dfx.diffeqsolve(dfx.ODETerm(l96), dfx.Tsit5(), 0.0, 0.5, None, yw, stepsize_controller=cw)
for mode in ("peer", "nccl"):
    p3 = dfx.prepare_sharded(dfx.ODETerm(l96), dfx.Tsit5(), 0.0, 0.5, None, yw, stepsize_controller=cw, gather=mode)
    o3 = p3()
    torch.cuda.synchronize()
    assert torch.equal(o3.y_final, r3.ys[:, 0]) and int(o3.stats["num_steps"]) == int(r3.stats["num_steps"].sum()), (rank, mode)
    p3.close()
dist.barrier()
if rank == 0:
    print(f"peer gather ok: world {world}, {n} trajectories, device + host inputs, peer == nccl == single-GPU; wide state (d = {D}) too")
dist.destroy_process_group()
