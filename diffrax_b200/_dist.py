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
import ctypes as C
import dataclasses
from typing import Any, Optional


@dataclasses.dataclass
class ShardedSolution:
    """Result of `sharded_diffeqsolve` on one rank.

    local     this rank's `Solution` (rows [lo, hi) of the global batch; host or device arrays like the inputs)
    y_final   [n_total, d] final states of the WHOLE batch, on this rank's device (NCCL all_gather), block order
    t_final   [n_total]
    stats     global totals: num_steps / num_accepted_steps / num_rejected_steps / num_failed (result neither successful nor
              event_occurred) / max_steps_per_trajectory
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


class _DevBuf:
    """A raw device allocation (dfx_peer_alloc / dfx_peer_open) exposed to torch through __cuda_array_interface__."""

    def __init__(self, ptr, nbytes):
        self.ptr, self.nbytes = int(ptr), int(nbytes)
        self.__cuda_array_interface__ = {"shape": (self.nbytes,), "typestr": "|u1", "data": (self.ptr, False), "version": 2}

    def tensor(self, device):
        return torch.as_tensor(self, device=device)


class _PeerGather:
    """The fused gather: every rank owns TWO global buffers [finals (N x d) | t_final (N)] in peer-mappable device memory
    (CUDA IPC), maps its peers' buffers, and hands all of them to the solve kernel, which stores each trajectory's final
    state into every rank's buffer (P2P stores over NVLink / NVSwitch) the moment the trajectory is finalised.  The transfer
    overlaps the solve; afterwards a 32-byte all_gather of the in-kernel totals is both the barrier and the statistics
    reduction.  The two buffers alternate between calls, so a rank that races ahead into the next solve writes the buffer its
    peers are NOT reading."""

    def __init__(self, n_total, d, dtype, device, group):
        from . import _lib
        self.L = _lib.lib()
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.n_total, self.d, self.dtype, self.device = n_total, d, dtype, device
        es = torch.empty((), dtype=dtype).element_size()
        self.nbytes = (n_total * (d + 1) * es + 255) // 256 * 256
        self.own, handles = [], []
        with torch.cuda.device(device):
            for _ in range(2):
                ptr, h = C.c_void_p(), C.create_string_buffer(64)
                _lib.check(self.L.dfx_peer_alloc(self.nbytes, C.byref(ptr), h))
                self.own.append(ptr.value)
                handles.append(h.raw)
            gathered = [None] * self.world
            dist.all_gather_object(gathered, handles, group=group)
            self.mapped = []          # [buffer][rank] -> device pointer valid here
            self._opened = []
            for b in range(2):
                row = []
                for r in range(self.world):
                    if r == self.rank:
                        row.append(self.own[b])
                    else:
                        q = C.c_void_p()
                        _lib.check(self.L.dfx_peer_open(gathered[r][b], C.byref(q)))
                        row.append(q.value)
                        self._opened.append(q.value)
                self.mapped.append(row)
        self.views = []
        for b in range(2):
            raw = _DevBuf(self.own[b], self.nbytes).tensor(device)
            y = raw[: n_total * d * es].view(dtype).view(n_total, d)
            t = raw[n_total * d * es: n_total * (d + 1) * es].view(dtype)
            self.views.append((y, t, raw))
        self.t_off = n_total * d * es
        self.turn = 0

    def bind(self, desc, lo, which):
        desc.n_peers = self.world
        desc.peer_row_offset = lo
        for r in range(self.world):
            desc.peer_y_final[r] = self.mapped[which][r]
            desc.peer_t_final[r] = self.mapped[which][r] + self.t_off

    def close(self):
        for q in self._opened:
            self.L.dfx_peer_close(q)
        self._opened = []
        for p_ in self.own:
            self.L.dfx_peer_free(p_)
        self.own = []


class ShardedSolve:
    """A prepared sharded solve (see `prepare_sharded`): every call runs this rank's shard and makes every rank see the final
    states of the WHOLE batch plus the global step statistics.

    gather="peer" (default on one node with world > 1): the all_gather of the finals is FUSED into the solve kernel - P2P
    stores into every rank's buffer over NVLink as trajectories finish (`_PeerGather`) - and the only collective left is a
    32-byte all_gather of the in-kernel totals, which doubles as the barrier.
    gather="nccl": ONE all_gather of a packed per-rank record [finals | t_final | 4 int64 totals] after the kernel; the kernel
    writes the finals and the totals straight into the record."""

    def __init__(self, plan, lo, hi, n_total, d, dtype, device, group, gather="nccl"):
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
        self.peer = None
        if gather == "peer" and self.world > 1 and device.type == "cuda":
            # every rank must end up in the same mode: agree on whether peer memory could be set up everywhere
            # (it cannot across nodes, or where CUDA IPC / P2P access is unavailable) and fall back to the NCCL record together
            try:
                peer, ok = _PeerGather(n_total, d, dtype, device, group), 1
            except Exception as e:  # noqa: BLE001
                peer, ok, self.peer_error = None, 0, f"{type(e).__name__}: {e}"
            flag = torch.tensor([ok], dtype=torch.int32, device=device)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
            if int(flag[0]) == 1:
                self.peer = peer
                self.tot_all = torch.empty(self.world * 4, dtype=torch.int64, device=device)
            elif peer is not None:
                peer.close()

    def __call__(self, throw: bool = True) -> ShardedSolution:
        # the block is always solved with throw=False: one rank raising before the collective would leave the others waiting;
        # `throw` is applied to the GLOBAL failure count afterwards, so every rank raises (or none does) - _integrate.py:1541-1542
        out = self.gather(self.solve_local(throw=False))
        if throw and int(out.stats["num_failed"]) > 0:
            raise RuntimeError(f"{int(out.stats['num_failed'])} of {self.n_total} trajectories failed (result codes in `.local.result`; "
                               "pass throw=False to inspect them)")
        return out

    def solve_local(self, throw: bool = True):
        """This rank's block: one C-ABI call; the finals and the totals land in the packed record (and, in peer mode, in
        every rank's global buffer)."""
        if self.peer is not None:
            self.peer.turn ^= 1
            self.peer.bind(self.plan.desc, self.lo, self.peer.turn)
        return self.plan(throw=throw)

    def gather(self, sol) -> ShardedSolution:
        """Make the global finals + statistics visible: a 32-byte all_gather (peer mode: barrier + totals) or the one
        all_gather of every rank's [finals | t_final | totals] record."""
        d, es = self.d, self._es
        if self.peer is not None:
            dist.all_gather_into_tensor(self.tot_all, self.totals, group=self.group)   # also orders every rank's P2P stores before the reads
            y, t, _ = self.peer.views[self.peer.turn]
            tot = self.tot_all.view(self.world, 4)
        elif self.world > 1:
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
        out = ShardedSolution(sol, self.lo, self.hi, self.n_total, y, t, stats)
        out._owner = self   # y_final / t_final may be views of this object's peer buffers: keep them alive with the result
        return out

    def close(self):
        """Unmap the peers' buffers and free this rank's (peer mode); safe to call more than once."""
        if getattr(self, "peer", None) is not None:
            self.peer.close()
            self.peer = None

    def __del__(self):
        try:
            self.close()
        except Exception:  # noqa: BLE001  (interpreter shutdown)
            pass


def prepare_sharded(terms, solver, t0, t1, dt0, y0, args=None, *, group=None, device=None, gather=None, **kw) -> ShardedSolve:
    """`prepare` for the sharded configuration.  `y0` (and per-trajectory `t0` / `t1`, Brownian keys) describe the GLOBAL
    batch of N trajectories and are the same on every rank; rank r owns the contiguous block `shard_range(N, r, world)`.
    Host inputs (NumPy / CPU tensors) go through the host-buffer entry of the C ABI, CUDA tensors through the device entry.
    `gather`: "peer" (default with NCCL on one node) fuses the gather of the finals into the solve kernel over NVLink peer
    memory; "nccl" distributes a packed per-rank record with one all_gather (see `ShardedSolve`).  Call `.close()` on the
    returned object when done with it (peer mode holds CUDA-IPC mappings)."""
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
    if gather is None:   # the fused peer gather on one node; DFX_GATHER=nccl selects the plain collective
        gather = os.environ.get("DFX_GATHER", "peer" if (world > 1 and dist.get_backend(group) == "nccl") else "nccl")
    if gather not in ("peer", "nccl"):
        raise ValueError("gather must be 'peer' or 'nccl'")
    sh = ShardedSolve.__new__(ShardedSolve)
    # two-phase construction: the record buffers must exist before `prepare` binds them as final_out
    ShardedSolve.__init__(sh, None, lo, hi, n_total, d, dtype, device, group, gather)
    sh.plan = _api.prepare(_slice_terms(terms, lo, hi), solver, _slice_rows(t0, lo, hi), _slice_rows(t1, lo, hi), dt0, y_loc,
                           None if args is None else args[lo:hi],   # per-trajectory parameters follow their trajectories
                           device=device.index if device.index is not None else 0, final_out=(sh.y_buf, sh.t_buf, sh.totals), **kw)
    return sh


def sharded_diffeqsolve(terms, solver, t0, t1, dt0, y0, args=None, *, throw: bool = True, group=None, device=None, **kw) -> ShardedSolution:
    """`diffeqsolve` for a global batch sharded over the ranks of a `torch.distributed` group (one process per GPU):
    trajectories are independent, so each rank integrates its block with no data-path collective; the final states of the
    whole batch reach every rank through the fused peer gather (or one all_gather), the global step statistics through a
    32-byte all_gather.  `throw` is collective: every rank raises if ANY trajectory of the global batch failed."""
    return prepare_sharded(terms, solver, t0, t1, dt0, y0, args, group=group, device=device, **kw)(throw=throw)
