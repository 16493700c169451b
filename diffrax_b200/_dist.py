"""Multi-GPU plumbing: one process per GPU, trajectories block-partitioned across ranks.

The path shards naturally (independent trajectories): there is NO data-path
collective.  NCCL (or gloo in the CPU tests) is used only after the kernel to gather final
states and to reduce the step statistics.
"""
from __future__ import annotations

import os
from typing import Tuple

import torch
import torch.distributed as dist


def shard_range(n_total: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous block [lo, hi) of trajectories owned by `rank`; sizes differ by at most 1."""
    base, rem = divmod(int(n_total), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def init_from_env(backend: str | None = None):
    """torchrun-style rendezvous (RANK / WORLD_SIZE / MASTER_ADDR / MASTER_PORT)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", str(rank)))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        kw = {}
        if backend == "nccl":
            kw["device_id"] = torch.device("cuda", local)
        dist.init_process_group(backend=backend, rank=rank, world_size=world, **kw)
    return rank, local, world


def gather_final_states(y_final: torch.Tensor, n_total: int) -> torch.Tensor:
    """all_gather of the ranks' [n_local, d] final states into the global [n_total, d] array
    (block partition order).  Shards may differ in length by one row, so they are padded."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return y_final
    world = dist.get_world_size()
    sizes = [shard_range(n_total, r, world) for r in range(world)]
    nmax = max(hi - lo for lo, hi in sizes)
    pad = torch.zeros((nmax, y_final.shape[1]), dtype=y_final.dtype, device=y_final.device)
    pad[: y_final.shape[0]] = y_final
    out = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(out, pad)
    return torch.cat([o[: hi - lo] for o, (lo, hi) in zip(out, sizes)], dim=0)


def reduce_stats(stats: dict) -> dict:
    """Sum of attempted / accepted / rejected steps and the max per-trajectory step count over all
    ranks; `num_failed` counts trajectories whose result != successful."""
    ns = stats["num_steps"].to(torch.int64)
    vec = torch.stack([ns.sum(), stats["num_accepted_steps"].to(torch.int64).sum(),
                       stats["num_rejected_steps"].to(torch.int64).sum(),
                       torch.as_tensor(int(stats.get("num_failed", 0)), device=ns.device, dtype=torch.int64)])
    mx = ns.max().reshape(1) if ns.numel() else torch.zeros(1, dtype=torch.int64, device=ns.device)
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(vec, op=dist.ReduceOp.SUM)
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
    return {"num_steps": int(vec[0]), "num_accepted_steps": int(vec[1]), "num_rejected_steps": int(vec[2]),
            "num_failed": int(vec[3]), "max_steps_per_trajectory": int(mx[0])}


# --------------------------------------------------------------------------------------
# Product-level sharded entry: the BASELINE configuration "N trajectories sharded across 1/2/4/8 B200"
# --------------------------------------------------------------------------------------
import dataclasses
from typing import Any, Optional


@dataclasses.dataclass
class ShardedSolution:
    """Result of `sharded_diffeqsolve` on one rank.

    local     this rank's `Solution` (rows [lo, hi) of the global batch; host or device arrays like the inputs)
    y_final   [n_total, d] final states of the WHOLE batch, on this rank's device (NCCL all_gather), block order
    t_final   [n_total]
    stats     global totals: num_steps / num_accepted_steps / num_rejected_steps / num_failed / max_steps_per_trajectory
    """
    local: Any
    lo: int
    hi: int
    n_total: int
    y_final: Optional[torch.Tensor]
    t_final: Optional[torch.Tensor]
    stats: dict


def _slice_rows(x, lo, hi):
    return x if x is None or not hasattr(x, "shape") or len(x.shape) == 0 else x[lo:hi]


def _slice_terms(terms, lo, hi):
    """The same term structure over trajectories [lo, hi): per-trajectory Brownian keys are sliced."""
    from . import _api
    if isinstance(terms, _api.MultiTerm):
        return _api.MultiTerm(*[_slice_terms(t, lo, hi) for t in terms.terms])
    if isinstance(terms, _api.ControlTerm):
        bm = terms.control
        keys = _api._as_keys(bm.key)[lo:hi]
        sub = _api.VirtualBrownianTree(bm.t0, bm.t1, bm.tol, bm.shape, keys, bm.levy_area, partitionable=bm.partitionable)
        return _api.ControlTerm(terms.vector_field, sub)
    return terms


class ShardedSolve:
    """A prepared sharded solve (see `prepare_sharded`): every call runs this rank's shard and then ONE collective - an
    all_gather of a packed per-rank record [finals | t_final | 4 int64 totals] - so the gather of the final states and the
    reduction of the statistics cost a single NCCL launch.  The solve kernel writes the finals AND
    the ensemble totals (reduced in-kernel) straight into the record; nothing is staged or reduced in between."""

    def __init__(self, plan, lo, hi, n_total, d, dtype, device, group):
        self.plan, self.lo, self.hi, self.n_total, self.d, self.group = plan, lo, hi, n_total, d, group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.spans = [shard_range(n_total, r, self.world) for r in range(self.world)]
        self.nmax = max(h - l for l, h in self.spans) if self.spans else 0
        es = torch.empty((), dtype=dtype).element_size()
        n = hi - lo
        self._off_t = self.nmax * d * es
        self._off_tot = (self.nmax * (d + 1) * es + 7) // 8 * 8
        self.rec = self._off_tot + 4 * 8                         # bytes of one rank's record
        self.dtype, self._es = dtype, es
        self.send = torch.zeros(self.rec, dtype=torch.uint8, device=device)
        self.recv = torch.empty(self.world * self.rec, dtype=torch.uint8, device=device) if self.world > 1 else self.send
        self.y_buf = self.send[: n * d * es].view(dtype).view(n, d)        # the kernel writes the finals straight into the record
        self.t_buf = self.send[self._off_t: self._off_t + n * es].view(dtype)
        self.totals = self.send[self._off_tot:].view(torch.int64)          # [sum steps, sum accepted, failed, max steps]

    def __call__(self, throw: bool = True) -> ShardedSolution:
        return self.gather(self.solve_local(throw=throw))

    def solve_local(self, throw: bool = True):
        """This rank's block: one C-ABI call; the finals and the totals land in the packed record."""
        return self.plan(throw=throw)

    def gather(self, sol) -> ShardedSolution:
        """The one collective of the path: all_gather of every rank's [finals | t_final | totals] record."""
        d, es = self.d, self._es
        if self.world > 1:
            dist.all_gather_into_tensor(self.recv, self.send, group=self.group)
            rec = self.recv.view(self.world, self.rec)
            y = torch.cat([rec[r, : (h - l) * d * es].view(self.dtype).view(h - l, d) for r, (l, h) in enumerate(self.spans)], 0)
            t = torch.cat([rec[r, self._off_t: self._off_t + (h - l) * es].view(self.dtype) for r, (l, h) in enumerate(self.spans)], 0)
            tot = rec[:, self._off_tot:].view(torch.int64)               # [world, 4]
        else:
            y, t, tot = self.y_buf, self.t_buf, self.totals[None]
        sums = tot[:, :3].sum(0)
        stats = {"num_steps": sums[0], "num_accepted_steps": sums[1], "num_rejected_steps": sums[0] - sums[1], "num_failed": sums[2],
                 "max_steps_per_trajectory": tot[:, 3].max()}               # 0-d device tensors: no host sync here
        return ShardedSolution(sol, self.lo, self.hi, self.n_total, y, t, stats)


def prepare_sharded(terms, solver, t0, t1, dt0, y0, args=None, *, group=None, device=None, **kw) -> ShardedSolve:
    """`prepare` for the sharded configuration.  `y0` (and per-trajectory `t0` / `t1`, Brownian keys) describe the GLOBAL
    batch of N trajectories and are the same on every rank; rank r owns the contiguous block `shard_range(N, r, world)`.
    Host inputs (NumPy / CPU tensors) go through the host-buffer entry of the C ABI, CUDA tensors through the device entry;
    either way the finals land in a device record that one NCCL all_gather distributes."""
    from . import _api
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    n_total = int(y0.shape[0])
    lo, hi = shard_range(n_total, rank, world)
    is_dev = isinstance(y0, torch.Tensor) and y0.is_cuda
    if device is None:
        device = y0.device if is_dev else torch.device("cuda", torch.cuda.current_device())
    device = torch.device(device)
    y_loc = y0[lo:hi]
    if isinstance(y_loc, torch.Tensor):
        y_loc = y_loc.contiguous()
    d = 1 if y_loc.ndim == 1 else int(y_loc.shape[1])
    import numpy as np
    dtype = y_loc.dtype if isinstance(y_loc, torch.Tensor) else getattr(torch, str(np.asarray(y_loc).dtype))
    if dtype not in (torch.float64, torch.float32):
        dtype = torch.float64
    sh = ShardedSolve.__new__(ShardedSolve)
    # two-phase construction: the record buffers must exist before `prepare` binds them as final_out
    ShardedSolve.__init__(sh, None, lo, hi, n_total, d, dtype, device, group)
    sh.plan = _api.prepare(_slice_terms(terms, lo, hi), solver, _slice_rows(t0, lo, hi), _slice_rows(t1, lo, hi), dt0, y_loc, args,
                           device=device.index if device.index is not None else 0, final_out=(sh.y_buf, sh.t_buf, sh.totals), **kw)
    return sh


def sharded_diffeqsolve(terms, solver, t0, t1, dt0, y0, args=None, *, throw: bool = True, group=None, device=None, **kw) -> ShardedSolution:
    """`diffeqsolve` for a global batch sharded over the ranks of a `torch.distributed` group (one process per GPU):
    trajectories are independent, so each rank integrates its block with no data-path collective, and ONE all_gather
    then gives every rank the final states of the whole batch plus the global step statistics."""
    return prepare_sharded(terms, solver, t0, t1, dt0, y0, args, group=group, device=device, **kw)(throw=throw)
