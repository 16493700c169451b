"""Host-side mirror of the reference interface: argument checking that the reference does eagerly
in Python (_integrate.py:1036-1045, 1143-1149, 1222-1233; constant.py:41-45; _saveat.py:40-48),
fail-loud behaviour without a GPU, and the multi-GPU sharding plumbing under gloo (world_size 2)."""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

import diffrax_b200 as dfx

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TERM = dfx.ODETerm(dfx.fields.Lorenz())
Y0 = np.ones((4, 3))
PID = dfx.PIDController(rtol=1e-6, atol=1e-6)


def test_argument_errors_match_reference():
    with pytest.raises(ValueError, match="dt0"):           # _integrate.py:1036-1045
        dfx.prepare(TERM, dfx.Dopri5(), 0.0, 1.0, -0.1, Y0, stepsize_controller=PID)
    with pytest.raises(ValueError, match="please pass a value for `dt0`"):   # constant.py:41-45
        dfx.prepare(TERM, dfx.Dopri5(), 0.0, 1.0, None, Y0)
    with pytest.raises(RuntimeError, match="between t0 and t1"):            # _integrate.py:1228-1232
        dfx.prepare(TERM, dfx.Dopri5(), 0.0, 1.0, 0.1, Y0, saveat=dfx.SaveAt(ts=[0.5, 1.5]))
    with pytest.raises(RuntimeError, match="increasing or decreasing"):     # 1223-1227
        dfx.prepare(TERM, dfx.Dopri5(), 0.0, 1.0, 0.1, Y0, saveat=dfx.SaveAt(ts=[0.5, 0.2, 0.7]))
    with pytest.raises(ValueError, match="nothing will be saved"):          # _saveat.py:40-48
        dfx.SaveAt()
    with pytest.raises(RuntimeError, match="error estimates"):              # pid.py:461-469
        dfx.prepare(TERM, dfx.Euler(), 0.0, 1.0, 0.1, Y0, stepsize_controller=PID)
    ou = dfx.fields.OrnsteinUhlenbeck()
    bm = dfx.VirtualBrownianTree(0.0, 1.0, 1e-3, (), dfx.random.split(dfx.random.key(0), 4))
    sde = dfx.MultiTerm(dfx.ODETerm(ou.drift), dfx.ControlTerm(ou.diffusion, bm))
    with pytest.raises(ValueError, match="Euler's method"):                 # _integrate.py:1143-1149
        dfx.prepare(sde, dfx.Euler(), 0.0, 1.0, 0.1, np.ones((4, 1)), stepsize_controller=PID)
    with pytest.raises(ValueError, match="strictly less"):                  # tree.py:281
        dfx.VirtualBrownianTree(1.0, 1.0, 1e-3, (), dfx.random.key(0))
    with pytest.raises(ValueError, match="state dimension"):
        dfx.prepare(TERM, dfx.Dopri5(), 0.0, 1.0, 0.1, np.ones((4, 2)))
    with pytest.raises(ValueError, match="ShARK"):
        dfx.prepare(TERM, dfx.ShARK(), 0.0, 1.0, 0.1, Y0)
    with pytest.raises(TypeError):
        dfx.ODETerm(lambda t, y, args: -y)   # Python callables are not device functors


def test_output_shapes_like_vmapped_reference():
    """test_vmap.py:27-125: ts[N,1] / ys[N,4,d] / ts[N,4096] for steps=True with the default max_steps."""
    p = dfx.prepare(TERM, dfx.Dopri5(), 0.0, 1.0, 0.1, Y0)
    s = p._solution
    assert s.ts.shape == (4, 1) and s.ys.shape == (4, 1, 3)
    s = dfx.prepare(TERM, dfx.Dopri5(), 0.0, 1.0, 0.1, Y0, saveat=dfx.SaveAt(ts=[0.1, 0.2, 0.3, 0.4]))._solution
    assert s.ts.shape == (4, 4) and s.ys.shape == (4, 4, 3)
    s = dfx.prepare(TERM, dfx.Dopri5(), 0.0, 1.0, 0.1, Y0, saveat=dfx.SaveAt(steps=True))._solution
    assert s.ts.shape == (4, 4096) and s.ys.shape == (4, 4096, 3)
    s = dfx.prepare(TERM, dfx.Dopri5(), 0.0, 1.0, 0.1, Y0, saveat=dfx.SaveAt(steps=2, t1=True), max_steps=101)._solution
    assert s.ts.shape == (4, 51)
    assert dfx.RESULTS.successful == 0       # test_saveat_solution.py:21


@pytest.mark.skipif(torch.cuda.is_available(), reason="only meaningful on a box without a GPU")
def test_fails_loudly_without_a_gpu():
    """No CPU / PyTorch fallback: the product path raises when there is no CUDA device."""
    with pytest.raises(RuntimeError, match="CUDA device"):
        dfx.diffeqsolve(TERM, dfx.Dopri5(), 0.0, 1.0, None, Y0, stepsize_controller=PID)


@pytest.mark.skipif(torch.cuda.is_available(), reason="only meaningful on a box without a GPU")
def test_pipelined_host_entry_fails_loudly_without_a_gpu():
    """The size class that dfx_ensemble_solve_host runs as one launch with chunked transfers (>= 256K trajectories,
    adaptive, SaveAt(t1)) is refused the same way - before any stream, allocation or polling loop is set up."""
    y0 = np.ones((1 << 18, 3))
    with pytest.raises(RuntimeError, match="CUDA device"):
        dfx.diffeqsolve(TERM, dfx.Dopri5(), 0.0, 1.0, None, y0, stepsize_controller=PID)


def test_product_never_imports_the_oracle():
    """The oracle is test infrastructure: nothing in the package, the headers, the baseline arm's helpers or tools/ touches it
    (the check script that does - fuzz_parity.py - lives under tests/)."""
    import re
    paths = []
    for top in ("diffrax_b200", "include", "tools"):
        for dirpath, _, files in os.walk(os.path.join(ROOT, top)):
            if os.path.basename(dirpath) != "user":        # (lib/user: generated plugin sources, build artefacts)
                paths += [os.path.join(dirpath, f) for f in files if f.endswith((".py", ".cu", ".cuh", ".h", ".sh"))]
    assert len(paths) > 30
    for path in paths:
        txt = open(path).read()
        assert not re.search(r"^\s*(import|from)\s+oracle\b", txt, re.M), path
        assert "liboracle" not in txt and 'include "../../oracle' not in txt, path


def test_shard_range_partitions():
    from diffrax_b200._dist import shard_range
    for n, w in ((1 << 20, 8), (1000, 3), (7, 8), (0, 2)):
        spans = [shard_range(n, r, w) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        sizes = [hi - lo for lo, hi in spans]
        assert max(sizes) - min(sizes) <= 1


_WORKER = r'''
import os, sys, torch
sys.path.insert(0, {root!r})
import torch.distributed as dist
from diffrax_b200 import _dist
rank, local, world = _dist.init_from_env("gloo")
n_total, d = 11, 3
lo, hi = _dist.shard_range(n_total, rank, world)
full = torch.arange(n_total * d, dtype=torch.float64).reshape(n_total, d)
got = _dist.gather_final_states(full[lo:hi].clone(), n_total)
assert torch.equal(got, full), (rank, got)
stats = {{"num_steps": torch.arange(lo, hi, dtype=torch.int32) + 10, "num_accepted_steps": torch.arange(lo, hi, dtype=torch.int32) + 5,
          "num_rejected_steps": torch.full((hi - lo,), 5, dtype=torch.int32), "num_failed": rank}}
r = _dist.reduce_stats(stats)
assert r["num_steps"] == sum(range(n_total)) + 10 * n_total, r
assert r["num_accepted_steps"] == sum(range(n_total)) + 5 * n_total and r["num_rejected_steps"] == 5 * n_total
assert r["num_failed"] == sum(range(world)) and r["max_steps_per_trajectory"] == n_total - 1 + 10
dist.barrier(); dist.destroy_process_group()
print("ok", rank)
'''


def test_gather_and_reduce_world_size_2_gloo(tmp_path):
    """N>1 path on CPU: two processes, gloo, 127.0.0.1 rendezvous."""
    script = tmp_path / "worker.py"
    script.write_text(_WORKER.format(root=ROOT))
    import socket
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), LOCAL_RANK=str(r), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1",
                   MASTER_PORT=str(port), CUDA_VISIBLE_DEVICES="")
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE,
                                      stderr=subprocess.STDOUT, text=True))
    outs = [p.communicate(timeout=180)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    assert all("ok" in o for o in outs)


_WORKER_SHARDED = r'''
import os, sys, types, torch
sys.path.insert(0, {root!r})
import torch.distributed as dist
from diffrax_b200 import _dist
rank, local, world = _dist.init_from_env("gloo")
n_total, d = 11, 3
lo, hi = _dist.shard_range(n_total, rank, world)
full_y = torch.arange(n_total * d, dtype=torch.float64).reshape(n_total, d) * 0.5
full_t = torch.arange(n_total, dtype=torch.float64) + 100.0
steps = torch.arange(n_total, dtype=torch.int32) + 10
sh = _dist.ShardedSolve(None, lo, hi, n_total, d, torch.float64, torch.device("cpu"), None)
def plan(throw=True):              # stands in for the CUDA solve: writes this rank's finals into the packed record
    sh.y_buf.copy_(full_y[lo:hi]); sh.t_buf.copy_(full_t[lo:hi])
    res = (torch.arange(lo, hi) % 4 == 0)
    sh.totals.copy_(torch.tensor([int(steps[lo:hi].sum()), int(steps[lo:hi].sum()) - 3 * (hi - lo), int(res.sum()), int(steps[lo:hi].max())]))
    return types.SimpleNamespace(stats={{"num_steps": steps[lo:hi]}}, result=res.to(torch.int32))
sh.plan = plan
for _ in range(2):                 # the record is reused across calls
    out = sh(throw=False)
    assert torch.equal(out.y_final, full_y) and torch.equal(out.t_final, full_t), (rank, out.y_final)
    assert int(out.stats["num_steps"]) == int(steps.sum()) and int(out.stats["num_accepted_steps"]) == int(steps.sum()) - 3 * n_total
    assert int(out.stats["num_rejected_steps"]) == 3 * n_total and int(out.stats["num_failed"]) == 3
    assert int(out.stats["max_steps_per_trajectory"]) == n_total - 1 + 10 and (out.lo, out.hi, out.n_total) == (lo, hi, n_total)
try:                               # throw is collective: the failures sit in rank 0's and rank 1's blocks, BOTH ranks raise after the gather
    sh(throw=True); raised = False
except RuntimeError as e:
    raised = "3 of 11 trajectories failed" in str(e)
assert raised, rank
dist.barrier(); dist.destroy_process_group()
print("ok", rank)
'''


def test_sharded_solve_record_world_size_2_gloo(tmp_path):
    """The sharded product entry's one collective (packed [finals | t_final | statistics] record, uneven shards) under gloo
    with two processes; the CUDA solve is replaced by a stub that writes the rank's rows."""
    script = tmp_path / "worker_sharded.py"
    script.write_text(_WORKER_SHARDED.format(root=ROOT))
    import socket
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), LOCAL_RANK=str(r), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1",
                   MASTER_PORT=str(port), CUDA_VISIBLE_DEVICES="")
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE,
                                      stderr=subprocess.STDOUT, text=True))
    outs = [p.communicate(timeout=180)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    assert all("ok" in o for o in outs)


def test_subsaveat_union_assembly():
    """SaveAt(subs=...): the rows of a leaf are cut out of the ONE solve over the union of the leaves' ts (`_MultiSolve`): t0 first,
    then the leaf's ts (inf where the trajectory stopped before them), and the final value right after the ts it reached."""
    import types
    from diffrax_b200._api import _MultiSolve, save_y
    inf = np.inf
    # union solve: t0 column + ts {1, 2, 3, 4}; trajectory 0 ran to the end (t = 5), trajectory 1 stopped at t = 2.5
    ts_u = np.array([[0.0, 1.0, 2.0, 3.0, 4.0], [0.0, 1.0, 2.0, inf, inf]])
    ys_u = np.where(np.isfinite(ts_u), ts_u * 10.0, inf)[..., None] * np.ones(2)
    u = types.SimpleNamespace(ts=ts_u, ys=ys_u, t_final=np.array([5.0, 2.5]), y_final=np.array([[50.0, 50.0], [25.0, 25.0]]))
    ts, ys = _MultiSolve._from_union(u, (True, True, [2, 4], save_y))           # leaf: t0, ts = {2, 4}, t1
    assert np.array_equal(ts, [[0.0, 2.0, 4.0, 5.0], [0.0, 2.0, 2.5, inf]])
    assert np.array_equal(ys[..., 0], [[0.0, 20.0, 40.0, 50.0], [0.0, 20.0, 25.0, inf]])
    ts, ys = _MultiSolve._from_union(u, (False, True, [], save_y))              # leaf: t1 only
    assert np.array_equal(ts, [[5.0], [2.5]]) and np.array_equal(ys[:, 0, 0], [50.0, 25.0])
    ts, ys = _MultiSolve._from_union(u, (False, False, [3], lambda t, y, args: y.sum(-1)))   # leaf: ts = {3}, fn
    assert np.array_equal(ts, [[3.0], [inf]]) and np.array_equal(ys, [[60.0], [inf]])
    # the planner: leaves without steps share a solve, a leaf with steps keeps its own; union columns follow the direction
    subs = [dfx.SubSaveAt(t1=True), dfx.SubSaveAt(t0=True, ts=[2.0, 1.0]), dfx.SubSaveAt(steps=True), dfx.SubSaveAt(ts=[2.5, 1.0])]
    p = dfx.prepare(TERM, dfx.Dopri5(), 3.0, 0.0, None, Y0, saveat=dfx.SaveAt(subs=subs), stepsize_controller=PID, max_steps=64)
    assert [k for k, _ in p.providers] == ["union", "union", "own", "union"]
    assert p.union.desc.n_save_ts == 3 and p.union.desc.save_t0 == 1 and p.union.desc.save_t1 == 0     # {2.5, 2, 1} backwards
    assert p.providers[1][1][2] == [2, 3] and p.providers[3][1][2] == [1, 3]
    with pytest.raises(RuntimeError, match="increasing or decreasing"):
        dfx.prepare(TERM, dfx.Dopri5(), 0.0, 3.0, None, Y0, saveat=dfx.SaveAt(subs=[dfx.SubSaveAt(ts=[1.0, 0.5, 2.0]), dfx.SubSaveAt(t1=True)]),
                    stepsize_controller=PID)
