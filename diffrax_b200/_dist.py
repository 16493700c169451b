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
