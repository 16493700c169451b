#!/usr/bin/env python
"""bench.py - headline benchmark of the ensemble integrator (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--workload c2|c1|c3|c4|c5_heun|c5_shark]
                    [--scaling strong|weak] [--no-extras]

metric   accepted RK steps / s across the ensemble (whole job, all ranks)
step     ONE pass of the hot path over one batch of synthetic input: a complete sharded `diffeqsolve` of the whole
         ensemble - each rank integrates its block of trajectories, then ONE NCCL all_gather distributes the final states
         and the step statistics; the collective is INSIDE the timed region
workload C2 (BASELINE.json configs[1]): Lorenz sigma=10 rho=28 beta=8/3, Dopri5, PIDController(rtol=atol=1e-8), dt0=None,
         t in [0,2], SaveAt(t1=True), fp64, 2^20 trajectories IN TOTAL, sharded across the N GPUs ("scaling": "strong");
         the weak-scaling figure (2^20 trajectories per GPU) rides along as the `weak` field when N > 1
value    device time (CUDA events on the launch stream around solve + gather, max over ranks), inputs resident in HBM
e2e      same metric through the public `diffrax_b200.prepare_sharded(...)()` call with HOST (pinned) buffers: H2D of
         this rank's y0 block, D2H of its ys/ts/stats/result, the all_gather, and a host read of the global statistics
roofline FP64 FMA pipe: achieved = attempted steps x 316 flop / device time of the ensemble kernel
         of ONE rank; peak = DFMA-chain microbenchmark measured live (MEASURED_PEAKS.json has no FP64 figure)
config_results  at N=1 the other BASELINE configs (C1, C3, C4, C5 Heun, C5 ShARK) are measured the same way after the
         headline and appended to the one JSON line, each with value / roofline / e2e / clocks / cpu_baseline
--impl reference  times the reference's CPU path: live Diffrax under jax.vmap on the JAX CPU backend when
         `baseline.probe()` finds one (kind "reference"), else the CPU restatement in oracle/ (kind "port") on all host
         cores; bounded sample per step.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FLOP_PER_ATTEMPTED_STEP = {"c2": 316, "c1": 228, "c3": 1536, "c3_1e-4": 1536}
DEFAULT_TRAJECTORIES = {"c1": 1024, "c2": 1 << 20, "c3": 1 << 18, "c3_1e-4": 1 << 18, "c4": 1 << 16, "c5_heun": 1 << 20, "c5_shark": 1 << 20}


def workload(name: str, n: int, seed_offset: int = 0):
    """Synthetic inputs of SURVEY.md §8d (deterministic NumPy default_rng)."""
    base = dict(dt0=None, dtype=np.float64, save_ts=None, controller="pid", levy_area=None, save_dense=False,
                max_steps=4096, rtol=0.0, atol=0.0, keys=None)
    if name == "c2":
        rng = np.random.default_rng(1 + seed_offset)
        y0 = np.stack([rng.uniform(-15, 15, n), rng.uniform(-20, 20, n), rng.uniform(5, 45, n)], 1)
        return dict(base, field="lorenz", params=[10.0, 28.0, 8.0 / 3.0], solver="dopri5", y0=y0, t0=0.0, t1=2.0,
                    rtol=1e-8, atol=1e-8, label="C2 Lorenz/Dopri5/PID(1e-8,1e-8)/fp64/t in [0,2]/SaveAt(t1)")
    if name == "c1":
        rng = np.random.default_rng(0 + seed_offset)
        y0 = rng.uniform(0.5, 2.0, (n, 2))
        return dict(base, field="lotka_volterra", params=[1.5, -1.0, -3.0, 1.0], solver="tsit5", y0=y0, t0=0.0, t1=10.0,
                    rtol=1e-6, atol=1e-6, save_ts=np.linspace(0, 10, 100),
                    label="C1 Lotka-Volterra/Tsit5/PID(1e-6,1e-6)/fp64/SaveAt(ts=100)")
    if name in ("c3", "c3_1e-4"):
        # BASELINE config 3: the Arenstorf initial condition perturbed by 1e-3 N(0,1), one period, dense output.  2^18
        # trajectories x max_steps x 520 B of dense buffers must fit 180 GB of HBM, so max_steps = 768 (105 GB); trajectories the
        # perturbation sends through a lunar near-collision need more steps than that and end as max_steps_reached - they are
        # counted in `failed_trajectories` and their accepted steps are real work.
        rng = np.random.default_rng(2 + seed_offset)
        # "c3_1e-4": the same with a 1e-4 perturbation, under which (almost) every trajectory completes the period within 768 steps
        sig = 1e-3 if name == "c3" else 1e-4
        y0 = np.array([0.994, 0.0, 0.0, -2.00158510637908252]) + sig * rng.standard_normal((n, 4))
        return dict(base, field="cr3bp", params=[0.012277471], solver="dopri8", y0=y0, t0=0.0, t1=17.0652165601579625,
                    rtol=1e-12, atol=1e-12, save_dense=True, max_steps=768,
                    label=f"C3 CR3BP(Arenstorf+{sig:g} N(0,1))/Dopri8/PID(1e-12,1e-12)/fp64/one period/SaveAt(dense), max_steps=768")
    if name == "c4":
        import diffrax_b200 as dfx
        mlp = dfx.fields.MLP.init(3, d=4, width=128)          # weights ~ U(+-1/sqrt(fan_in)), seed 3
        rng = np.random.default_rng(30 + seed_offset)
        y0 = rng.standard_normal((n, 4)).astype(np.float32)
        return dict(base, field="mlp", params=mlp.oracle_params(), mlp=mlp, solver="tsit5", dtype=np.float32, y0=y0,
                    t0=0.0, t1=10.0, rtol=1e-3, atol=1e-6,
                    label="C4 neural ODE MLP(4->128->128->4, softplus, tanh)/Tsit5/PID(1e-3,1e-6)/fp32/t in [0,10]/SaveAt(t1)")
    if name in ("c5_heun", "c5_shark"):
        import diffrax_b200 as dfx
        keys = dfx.random.split(dfx.random.key(seed_offset), n)
        sh = name == "c5_shark"
        return dict(base, field="ou", params=[1.0, 0.0, 0.5], solver="shark" if sh else "heun", dtype=np.float32,
                    y0=np.ones((n, 1), np.float32), t0=0.0, t1=1.0, dt0=2.0 ** -6, controller="constant",
                    levy_area="stla" if sh else "bi", keys=keys, bm_tol=2.0 ** -8,
                    label=f"C5 OU/{'ShARK+SpaceTimeLevyArea' if sh else 'Heun+BrownianIncrement'}/VirtualBrownianTree(tol=2^-8)"
                          "/ConstantStepSize(2^-6)/fp32/SaveAt(t1)")
    raise ValueError(name)


def _ours_objects(w, dev=None):
    import torch
    import diffrax_b200 as dfx
    F = {"lorenz": dfx.fields.Lorenz, "lotka_volterra": dfx.fields.LotkaVolterra, "cr3bp": dfx.fields.CR3BP,
         "ou": dfx.fields.OrnsteinUhlenbeck, "mlp": None}[w["field"]]
    S = {"dopri5": dfx.Dopri5, "tsit5": dfx.Tsit5, "dopri8": dfx.Dopri8, "heun": dfx.Heun, "shark": dfx.ShARK}[w["solver"]]
    field = w["mlp"] if w["field"] == "mlp" else F(*w["params"])
    if w["levy_area"]:
        if dev is None:  # host path: the keys are an input of every step, so they sit in pinned memory like y0
            keys = torch.from_numpy(w["keys"].view(np.int32).copy())
            keys = keys.pin_memory() if torch.cuda.is_available() else keys
        else:
            keys = torch.tensor(w["keys"].view(np.int32), device=dev)
        lv = dfx.BrownianIncrement if w["levy_area"] == "bi" else dfx.SpaceTimeLevyArea
        bm = dfx.VirtualBrownianTree(0.0, 1.0, w["bm_tol"], (), keys, lv)
        term = dfx.MultiTerm(dfx.ODETerm(field.drift), dfx.ControlTerm(field.diffusion, bm))
    else:
        term = dfx.ODETerm(field)
    ctrl = dfx.PIDController(rtol=w["rtol"], atol=w["atol"]) if w["controller"] == "pid" else dfx.ConstantStepSize()
    if w["save_dense"]:
        saveat = dfx.SaveAt(dense=True)
    else:
        saveat = dfx.SaveAt(t1=True) if w["save_ts"] is None else dfx.SaveAt(ts=w["save_ts"])
    return dfx, term, S(), ctrl, saveat


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md)."""

    def __init__(self, index: int):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "20"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
                for nm, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def ncu_traffic(workload: str):
    """dram__bytes_read.sum + dram__bytes_write.sum of the dominant kernel from the committed `ncu --set full`
    summary (profiles/r0N_<workload>_*_ncu_full.txt, newest round first), in bytes per launch; None when no capture."""
    import glob
    import re
    for f in sorted(glob.glob(os.path.join(ROOT, "profiles", f"r0*_{workload}_*ncu_full.txt")), reverse=True):
        tot, units = 0.0, {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        for m in re.finditer(r"dram__bytes_(read|write)\.sum \[(\w+)\] = ([0-9.]+)", open(f).read()):
            tot += float(m.group(3)) * units.get(m.group(2), 1.0)
        if tot > 0:
            return tot
    return None


def cpu_port_rate(w, sample: int, threads: int = 0):
    """accepted steps / s of the oracle (CPU restatement) on `sample` trajectories."""
    import oracle
    y0 = w["y0"][:sample]
    t = time.perf_counter()
    o = oracle.solve(w["field"], y0, w["t0"], w["t1"], w["dt0"], solver=w["solver"], params=w["params"],
                     rtol=w["rtol"], atol=w["atol"], dtype=w["dtype"], save_t1=w["save_ts"] is None and not w["save_dense"],
                     save_ts=w["save_ts"], num_threads=threads, controller=w["controller"], levy_area=w["levy_area"],
                     keys=None if w["keys"] is None else w["keys"][:sample], bm_tol=w.get("bm_tol", 1e-3),
                     save_dense=w["save_dense"], max_steps=w["max_steps"])
    dt = time.perf_counter() - t
    return float(o["stats"][:, 1].sum()) / dt, dt, oracle.hw_threads() if threads == 0 else threads


_LIVE = {}


def live_reference_rate(w, sample: int):
    """accepted steps / s of LIVE Diffrax (jax.jit(jax.vmap(diffeqsolve)) on the JAX CPU backend, the pattern of
    benchmarks/lotka_volterra.py:55-59) on `sample` trajectories, compiled once per (workload, sample)."""
    import baseline
    case = dict(w, y0=w["y0"][:sample], keys=None if w["keys"] is None else w["keys"][:sample],
                save_t1=w["save_ts"] is None and not w["save_dense"])
    case.pop("mlp", None); case.pop("label", None)
    key = (w["label"], sample)
    if key not in _LIVE:
        fn, a = baseline.compiled(case)
        import jax
        jax.block_until_ready(fn(*a).stats["num_accepted_steps"])      # compile + warm-up
        _LIVE[key] = (fn, a)
    fn, a = _LIVE[key]
    import jax
    t = time.perf_counter()
    sol = fn(*a)
    acc = jax.block_until_ready(sol.stats["num_accepted_steps"])
    dt = time.perf_counter() - t
    return float(np.asarray(acc).sum()) / dt, dt, os.cpu_count() or 1


def cpu_reference_rate(w, sample: int):
    """(rate, seconds, cores, kind): the reference itself when importable, else the oracle port."""
    import baseline
    ok, why = baseline.probe()
    if ok:
        try:
            return (*live_reference_rate(w, sample), "reference", why)
        except Exception as e:  # noqa: BLE001  (a JAX twin missing for this case, an API drift ...): say so and fall back
            why = f"live Diffrax failed on this workload ({type(e).__name__}: {e})"
    return (*cpu_port_rate(w, sample), "port", why)


# bounded CPU samples (trajectories) sized for a few seconds of host work per config
CPU_SAMPLE = {"c1": 1024, "c2": 1 << 19, "c3": 1 << 13, "c3_1e-4": 1 << 13, "c4": 1 << 14, "c5_heun": 1 << 18, "c5_shark": 1 << 18}


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path on the box's host cores - live Diffrax when
    `import jax, diffrax` works (baseline.probe()), else the oracle port - on the arm's own config / metric / unit."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sample = min(args.ref_sample or CPU_SAMPLE[args.workload], DEFAULT_TRAJECTORIES[args.workload])
    w = workload(args.workload, sample)
    for _ in range(min(args.warmup, 1)):
        cpu_reference_rate(w, min(sample, 8192))
    rates, times = [], []
    cores, kind, why = 1, "port", ""
    for _ in range(args.steps):
        r, dt, cores, kind, why = cpu_reference_rate(w, sample)
        rates.append(r); times.append(dt)
    total_t = sum(times)
    value = float(np.sum(np.array(rates) * np.array(times)) / total_t)
    what = ("live Diffrax, jax.jit(jax.vmap(diffeqsolve)) on the JAX CPU backend" if kind == "reference"
            else "oracle (C restatement of the reference; live Diffrax unavailable: " + why + ")")
    smp = f"{sample} of the workload's trajectories per step (same seed/config), all host cores; {what}"
    line = {"impl": "reference", "metric": "accepted_rk_steps_per_s", "value": value, "unit": "steps/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total_t / args.steps,
            "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
            "dtype": "f64" if w["dtype"] == np.float64 else "f32", "data": "synthetic",
            "config": {"workload": w["label"], "trajectories_per_step": sample},
            "cpu_baseline": {"value": value, "unit": "steps/s", "cores": cores, "kind": kind, "sample": smp},
            "e2e": {"value": value, "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def _c5_blocks(name):
    """threefry blocks per step the descent-cached algorithm executes on the C5 time grid (dt = 2^-6, 8 levels)."""
    def walk(r, depth=8):
        s_, bits = 0.0, []
        for lv in range(depth):
            t_ = s_ + 2.0 ** -(lv + 1)
            right = r > t_
            bits.append(right)
            s_ = t_ if right else s_
        return bits
    nsteps_grid, lv_total, prev = 64, 0, walk(0.0)
    for k_ in range(1, nsteps_grid + 1):
        cur = walk(k_ / 64.0)
        common = next((i for i in range(8) if cur[i] != prev[i]), 8)
        lv_total += 8 - common
        prev = cur
    per_level, leaf = {"c5_heun": (3, 1), "c5_shark": (7, 4)}[name]
    return per_level * lv_total / nsteps_grid + leaf, {"c5_heun": 3 * 8 + 4, "c5_shark": 7 * 8 + 9}[name]


def _roofline(name, L, _lib, local, dev, att_per_gpu, acc_per_gpu, ms_per_step, out_bytes):
    import torch
    sec = ms_per_step * 1e-3
    flop = FLOP_PER_ATTEMPTED_STEP.get(name)
    if flop is not None:
        peak = float(L.dfx_measure_fma_peak(_lib.F64, local))  # TFLOP/s, live DFMA-chain microbenchmark
        achieved = att_per_gpu * flop / sec / 1e12  # per GPU: the kernel of ONE rank
        roof = {"bound": "fma_fp64", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                "frac": achieved / peak if peak > 0 else None, "traffic": None,
                "peak_source": "dfx_measure_fma_peak: 8 independent DFMA chains/thread, 8 CTAs x 256 thr/SM, "
                               "measured in this run (MEASURED_PEAKS.json holds no FP64 FMA figure)",
                "flop_per_attempted_step": flop}
        if name.startswith("c3"):   # dense output: 8 (16 d + 1) = 520 B per accepted step (algorithmic) + the +inf padding of the layout
            alg = acc_per_gpu * 520.0
            roof["hbm_write"] = {"algorithmic_bytes_per_launch": alg, "layout_bytes_per_launch": float(out_bytes),
                                 "achieved_GBps_algorithmic": alg / sec / 1e9, "achieved_GBps_layout": out_bytes / sec / 1e9,
                                 "peak_GBps": 6546.6, "peak_source": "MEASURED_PEAKS.json hbm_gbs (copy, read+write)"}
    elif name == "c4":
        # MLP field: 6 evaluations x 2 (128 d + 128^2 + 128 d) = 208 896 flop per attempted step;
        # 196 608 of them are the 128x128 hidden layer that runs on tcgen05 (issued 3x for 3xTF32, and the kernel
        # evaluates 7 stages per step: stage 0 is recomputed instead of carried, value-identical).
        flop = 208896
        a = torch.randn(8192, 8192, device=dev); b = torch.randn(8192, 8192, device=dev)
        torch.backends.cuda.matmul.allow_tf32 = True
        best = 1e9
        for _ in range(6):
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record(); torch.matmul(a, b); e1.record(); torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        del a, b
        peak = 2 * 8192.0 ** 3 / (best * 1e-3) / 1e12
        achieved = att_per_gpu * flop / sec / 1e12
        issued = att_per_gpu * 7 * 3 * 2 * 128 * 128 / sec / 1e12
        roof = {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                "frac": achieved / peak if peak > 0 else None, "traffic": None,
                "peak_source": "cuBLAS TF32 GEMM 8192^3 (torch.matmul, allow_tf32) measured in this run; "
                               "MEASURED_PEAKS.json has bf16 only",
                "flop_per_attempted_step": flop, "issued_tensor_tflops": issued,
                "note": "whole-solve average incl. the CUDA-core layers, softplus/tanh and the RK/PID algebra; "
                        "tensor-pipe utilisation of the kernel: profiles/ ncu summary"}
    else:
        # C5: threefry blocks on the INT32 ALU.  A query W(r) walks L = 8 tree levels (BI 3 blocks/level, STLA 7) and
        # draws the leaf bridge (BI 1 block, STLA 4); the step's other end point is the previous step's query, and the
        # descent cache (csrc/vbt.cuh) resumes each walk at the first level where it leaves the previous one - so the
        # ALGORITHM executes, per step, only the levels below the common prefix.  Counted exactly for this time grid.
        blocks, blocks_nocache = _c5_blocks(name)
        ops = blocks * 77.0          # 20 x (add, rotate, xor) + 17 injection adds per block
        peak = float(L.dfx_measure_int_peak(local))
        achieved = att_per_gpu * ops / sec / 1e12
        roof = {"bound": "int32_alu", "achieved": achieved, "peak": peak, "unit": "Tera-op/s",
                "frac": achieved / peak if peak > 0 else None, "traffic": None,
                "peak_source": "dfx_measure_int_peak: add/rotate/xor chains, measured in this run",
                "threefry_blocks_per_step": blocks, "threefry_blocks_per_step_without_descent_cache": blocks_nocache}
    roof["traffic"] = ncu_traffic(name)   # (a capture of this very workload, or none)
    return roof


def measure(name, args, rank, local, world, dev, strong: bool, with_cpu: bool, steps=None):
    """One workload through the sharded product entry: device-timed `value`, host-buffer `e2e`, roofline, clocks.
    strong: the workload's trajectory count is the TOTAL over all ranks; else it is the count PER GPU (weak)."""
    import torch
    import torch.distributed as dist
    import diffrax_b200 as dfx
    from diffrax_b200 import _lib
    L = _lib.lib()
    n_cfg = args.trajectories or DEFAULT_TRAJECTORIES[name]
    n_total = n_cfg if strong else n_cfg * world
    # the GLOBAL batch, identical on every rank (seeded); each rank integrates its contiguous block
    w = workload(name, n_total)
    _, term_d, solver, ctrl, saveat = _ours_objects(w, dev)
    y0_dev = torch.tensor(w["y0"], device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2

    # prepared sharded call: descriptor + output buffers + gather record built once; each step = ONE C-ABI call + ONE collective
    plan = dfx.prepare_sharded(term_d, solver, w["t0"], w["t1"], w["dt0"], y0_dev, saveat=saveat, stepsize_controller=ctrl,
                               max_steps=w["max_steps"])
    n_local = plan.hi - plan.lo

    def _tensors(x):
        if isinstance(x, torch.Tensor):
            yield x
        elif isinstance(x, dict):
            for v in x.values():
                yield from _tensors(v)
    out_bytes = sum(int(t.numel() * t.element_size()) for k in plan.plan._keep for t in _tensors(k) if t.is_cuda) \
        - int(y0_dev[plan.lo:plan.hi].numel() * y0_dev.element_size())
    do_e2e = out_bytes < (2 << 30)      # C3's ~100 GB of dense output is not staged through the host

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up (also pays cudaMallocAsync pool growth, module load, NCCL channel set-up) ----
    for _ in range(max(args.warmup, 3)):
        sol = plan(throw=False)
    torch.cuda.synchronize()
    acc = int(sol.stats["num_accepted_steps"]); att = int(sol.stats["num_steps"]); failed = int(sol.stats["num_failed"])
    loc = sol.local
    att_local = int(loc.stats["num_steps"].sum()); acc_local = int(loc.stats["num_accepted_steps"].sum())

    # ---- timed: K steps, per-step CUDA events on the launch stream, L2 flushed between steps ----
    K = steps or args.steps
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); plan(throw=False); e1.record(); torch.cuda.synchronize()
    est_ms = max(e0.elapsed_time(e1), 1e-3)
    if world > 1:                               # every rank must run the SAME number of steps (each one is a collective)
        em = torch.tensor([est_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(em, op=dist.ReduceOp.MAX)
        est_ms = float(em[0])
    if est_ms * K < 400.0:                      # short workloads: run long enough for the 20 ms clock sampler to see them
        K = int(math.ceil(400.0 / est_ms))
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.05)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    L.dfx_reset_launch_count()
    barrier()
    wall0 = time.perf_counter()
    for (a0, a1), (k0, k1) in zip(ev, kev):
        flush.fill_(1)  # L2 flush, outside the event pair
        a0.record()
        k0.record()
        loc_sol = plan.solve_local(throw=False)   # this rank's kernel ...
        k1.record()
        sol = plan.gather(loc_sol)                # ... and the one collective: finals + statistics of every rank
        a1.record()
    barrier()
    wall = time.perf_counter() - wall0
    launches = int(L.dfx_launch_count())
    clocks = sampler.stop() if rank == 0 else None
    dev_ms = sum(a.elapsed_time(b) for a, b in ev)     # solve + gather
    ker_ms = sum(a.elapsed_time(b) for a, b in kev)    # the ensemble kernel of this rank alone

    # ---- e2e: K steps through the host-buffer entry (H2D + D2H inside), gather + host read of the statistics ----
    e2e_s, e2e_mean_s, h2d, d2h = float("nan"), float("nan"), 0, 0
    if do_e2e:
        _, term_h, _, _, _ = _ours_objects(w, None)
        y0_host = torch.tensor(w["y0"]).pin_memory()
        hplan = dfx.prepare_sharded(term_h, solver, w["t0"], w["t1"], w["dt0"], y0_host, saveat=saveat, stepsize_controller=ctrl,
                                    max_steps=w["max_steps"], device=dev)
        for _ in range(2):
            sh = hplan(throw=False)
            _ = int(sh.stats["num_accepted_steps"])
        barrier()
        per_step = []
        Ke = min(K, max(args.steps, int(math.ceil(400.0 / max(est_ms, 0.2)))))
        t_all = time.perf_counter()
        for _ in range(Ke):
            t0 = time.perf_counter()
            sh = hplan(throw=False)
            _ = int(sh.stats["num_accepted_steps"])  # the step's result (global statistics) read on the host
            per_step.append(time.perf_counter() - t0)
        barrier()
        e2e_mean_s = (time.perf_counter() - t_all) / Ke
        # per-step MEDIAN (one scheduler hiccup of the host process would otherwise decide the figure); mean kept next to it
        e2e_s = float(np.median(per_step))
        hl = sh.local
        y_blk = y0_host[hplan.lo:hplan.hi]
        h2d = y_blk.numel() * y_blk.element_size() + (0 if w["keys"] is None else n_local * 8)
        d2h = sum(int(t.numel() * t.element_size()) for t in (hl.ts, hl.ys, hl.result)) + 3 * int(hl.stats["num_steps"].numel()) * 4 + 8
        hplan.close()
        del hplan, y0_host

    # ---- max over ranks ----
    t_max = torch.tensor([dev_ms, ker_ms, (e2e_s if do_e2e else 0.0) * 1e3, (e2e_mean_s if do_e2e else 0.0) * 1e3, wall * 1e3],
                         dtype=torch.float64, device=dev)
    cnt = torch.tensor([att_local, acc_local], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_max, op=dist.ReduceOp.MAX)
        dist.all_reduce(cnt, op=dist.ReduceOp.MAX)      # the busiest rank's kernel is the one the roofline describes
    dev_ms_max, ker_ms_max, e2e_ms_max, e2e_mean_ms_max, wall_ms_max = (float(x) for x in t_max)
    res = None
    if rank == 0:
        ms_per_step = dev_ms_max / K
        roof = _roofline(name, L, _lib, local, dev, float(cnt[0]), float(cnt[1]), ker_ms_max / K, out_bytes)
        res = {
            "metric": "accepted_rk_steps_per_s", "value": acc / (ms_per_step * 1e-3), "unit": "steps/s", "n_gpus": world,
            "steps": K, "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "kernel_ms_per_step": ker_ms_max / K,
            "higher_is_better": True, "scaling": "strong" if strong else "weak", "vs_baseline": None,
            "dtype": "f64" if w["dtype"] == np.float64 else "f32", "data": "synthetic",
            "config": {"workload": w["label"], "trajectories_total": n_total, "trajectories_per_gpu": n_local,
                       "accepted_steps_per_solve": acc, "attempted_steps_per_solve": att, "failed_trajectories": failed,
                       "l2": "flushed between timed steps (256 MiB write outside the per-step CUDA-event pairs)",
                       "parallelism": (f"trajectory-sharded x{world}; gather of the finals "
                                       + ("FUSED into the solve kernel (P2P stores into every rank's buffer over NVLink peer memory) + a 32-byte all_gather "
                                          "of the in-kernel totals" if plan.peer is not None else "by ONE all_gather of [finals | t_final | statistics]")
                                       + ", inside the timed region") if world > 1 else "single GPU (no collective)"},
            "e2e": ({"value": acc / (e2e_ms_max * 1e-3), "unit": "steps/s", "h2d_bytes_per_step": h2d,
                     "d2h_bytes_per_step": d2h, "ms_per_step": e2e_ms_max, "statistic": "median of the per-step wall times, max over ranks",
                     "mean_ms_per_step": e2e_mean_ms_max, "call": "diffrax_b200.prepare_sharded(host buffers)(): H2D, kernel, D2H, all_gather, host read of the statistics"}
                    if do_e2e else
                    {"value": None, "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0,
                     "note": "outputs exceed 2 GiB; not staged through the host"}),
            "gpu_launches": launches,
            "roofline": roof,
            "clocks": clocks, "wall_ms_per_step_incl_flush_and_host": wall_ms_max / K,
        }
        if with_cpu:
            cpu_sample = min(args.cpu_sample or CPU_SAMPLE[name], n_total)
            rate, cpu_dt, cores, kind, why = cpu_reference_rate(w, cpu_sample)
            res["cpu_baseline"] = {"value": rate, "unit": "steps/s", "cores": cores, "kind": kind,
                                   "sample": f"first {cpu_sample} trajectories of the batch, "
                                             + ("live Diffrax (jax.vmap, JAX CPU)" if kind == "reference" else f"oracle C port (live Diffrax: {why})")
                                             + f", {cpu_dt:.2f} s"}
    plan.close()
    del plan, y0_dev, flush, sol
    torch.cuda.empty_cache()
    return res


def parity_report(dev):
    """the run also emits a small parity report - this arm against live Diffrax when the probe finds
    it, else against the oracle: Brownian increments bit-exact, accepted-step counts, saved states (C2 slice, fp64)."""
    import torch
    import diffrax_b200 as dfx
    import oracle
    import baseline
    n = 2048
    w = workload("c2", n)
    sol = dfx.diffeqsolve(dfx.ODETerm(dfx.fields.Lorenz(*w["params"])), dfx.Dopri5(), 0.0, 2.0, None, torch.tensor(w["y0"], device=dev),
                          stepsize_controller=dfx.PIDController(rtol=1e-8, atol=1e-8))
    ok, why = baseline.probe()
    ref = None
    if ok:
        try:
            ref, against = baseline.solve(dict(w, save_t1=True)), "live Diffrax (" + why + ")"
        except Exception as e:  # noqa: BLE001
            why = f"live Diffrax failed: {type(e).__name__}: {e}"
    if ref is None:
        ref = oracle.solve("lorenz", w["y0"], 0.0, 2.0, None, solver="dopri5", params=w["params"], rtol=1e-8, atol=1e-8)
        against = "oracle (C restatement; live Diffrax unavailable: " + why + ")"
    ys = sol.ys.cpu().numpy()
    rys = np.asarray(ref["ys"]).reshape(ys.shape)
    err = np.abs(ys - rys) / (np.abs(rys) + 1e-3 * np.abs(rys).max())
    dacc = np.abs(sol.stats["num_accepted_steps"].cpu().numpy() - np.asarray(ref["stats"])[:, 1])
    keys = dfx.random.split(dfx.random.key(0), 4096)
    kd = torch.tensor(keys.view(np.int32), device=dev)
    bit = True
    for lv, cls in (("bi", dfx.BrownianIncrement), ("stla", dfx.SpaceTimeLevyArea)):
        bm = dfx.VirtualBrownianTree(0.0, 1.0, 2.0 ** -8, (), kd, cls)
        W, H = bm.evaluate(torch.full((4096,), 0.25, device=dev), torch.full((4096,), 0.625, device=dev), use_levy=True)
        Wo, Ho = oracle.vbt_evaluate(keys, 0.25, 0.625, tol=2.0 ** -8, levy_area=lv, dtype=np.float32)
        bit = bit and np.array_equal(W.cpu().numpy(), Wo) and (lv == "bi" or np.array_equal(H.cpu().numpy(), Ho))
    return {"against": against, "c2_slice_trajectories": n, "max_rel_state_err": float(err.max()),
            "frac_within_1e-10": float((err.max(axis=(1, 2)) < 1e-10).mean()), "max_abs_accepted_step_diff": int(dacc.max()),
            "brownian_increments_bit_exact_vs_oracle": bool(bit), "brownian_sample": "4096 keys, W (and H) of [0.25, 0.625], tol 2^-8, fp32"}


def run_ours(args):
    import gc
    import torch
    import torch.distributed as dist
    from diffrax_b200 import _dist

    rank, local, world = _dist.init_from_env()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU path)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    strong = args.scaling == "strong"
    line = measure(args.workload, args, rank, local, world, dev, strong, with_cpu=True)
    if world > 1 and strong and not args.no_extras:
        gc.collect(); torch.cuda.empty_cache()
        wk = measure(args.workload, args, rank, local, world, dev, False, with_cpu=False)
        if rank == 0:
            line["weak"] = {k: wk[k] for k in ("value", "unit", "ms_per_step", "kernel_ms_per_step", "steps")}
            line["weak"].update(trajectories_per_gpu=wk["config"]["trajectories_per_gpu"], trajectories_total=wk["config"]["trajectories_total"],
                                e2e_value=wk["e2e"]["value"], e2e_ms_per_step=wk["e2e"].get("ms_per_step"))
    if rank == 0 and args.workload == "c2" and not args.trajectories:
        try:
            line["parity"] = parity_report(dev)
        except Exception as e:  # noqa: BLE001
            line["parity"] = {"error": f"{type(e).__name__}: {str(e)[:200]}"}
    if world == 1 and not args.no_extras and args.workload == "c2" and not args.trajectories:
        # the other BASELINE configs, same method, appended to the one JSON line
        extras = {}
        for name in ("c1", "c3", "c3_1e-4", "c4", "c5_heun", "c5_shark"):
            gc.collect(); torch.cuda.empty_cache()        # the previous config's buffers (C3: 105 GB) go back to the driver first
            try:
                extras[name] = measure(name, args, rank, local, world, dev, True, with_cpu=True, steps=min(args.steps, 10))
            except Exception as e:  # noqa: BLE001  (never lose the headline to an extra)
                extras[name] = {"error": f"{type(e).__name__}: {str(e)[:300]}"}
        line["config_results"] = extras
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def run_published_jump_step(args):
    """The ONE timing the reference publishes for this path (benchmarks/jump_step_timing.py:17-116):
    100 vmapped Heun SDE solves, dy = -0.2 y dt + dW, VirtualBrownianTree(0, 5, 2^-5, ()), dt0 = 0.5,
    ClipStepSizeController(PIDController(rtol=0, atol=1e-3, dtmin=2^-9, dtmax=1, pcoeff=0.3, icoeff=0.7),
    step_ts=linspace(0, 5, 129)), SaveAt(ts=step_ts), fp32 (the script does not enable x64);
    wall time of 3 back-to-back runs, best of 20.  Published: 0.23506 s on unstated hardware."""
    import torch
    import diffrax_b200 as dfx
    dev = torch.device("cuda", 0)
    keys = dfx.random.split(dfx.random.key(0), 100)
    step_ts = np.linspace(0, 5, 129).astype(np.float32)
    ou = dfx.fields.OrnsteinUhlenbeck(0.2, 0.0, 1.0)           # drift -0.2 y, diffusion 1
    bm = dfx.VirtualBrownianTree(0, 5, 2.0 ** -5, (), torch.tensor(keys.view(np.int32), device=dev))
    term = dfx.MultiTerm(dfx.ODETerm(ou.drift), dfx.ControlTerm(ou.diffusion, bm))
    ctrl = dfx.ClipStepSizeController(dfx.PIDController(rtol=0, atol=1e-3, dtmin=2.0 ** -9, dtmax=1.0, pcoeff=0.3, icoeff=0.7),
                                      step_ts=step_ts)
    y0 = torch.ones(100, 1, dtype=torch.float32, device=dev)

    def one():
        sol = dfx.diffeqsolve(term, dfx.Heun(), 0.0, 5.0, 0.5, y0, saveat=dfx.SaveAt(ts=step_ts), stepsize_controller=ctrl)
        return sol

    for _ in range(max(args.warmup, 3)):
        sol = one()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(20):
        t = time.perf_counter()
        for _ in range(3):
            sol = one()
        torch.cuda.synchronize()                                # == jax.block_until_ready
        best = min(best, time.perf_counter() - t)
    acc = int(sol.stats["num_accepted_steps"].sum())
    # the same three runs through a prepared solve (arguments checked and buffers allocated once - the analogue of calling
    # the reference's jit-compiled function), and the device time of one solve
    call = dfx.prepare(term, dfx.Heun(), 0.0, 5.0, 0.5, y0, saveat=dfx.SaveAt(ts=step_ts), stepsize_controller=ctrl)
    call()
    torch.cuda.synchronize()
    best_prepared = 1e9
    for _ in range(20):
        t = time.perf_counter()
        for _ in range(3):
            call()
        torch.cuda.synchronize()
        best_prepared = min(best_prepared, time.perf_counter() - t)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); call(); e1.record(); torch.cuda.synchronize()
    device_ms = e0.elapsed_time(e1)
    line = {"metric": "jump_step_timing_3_runs_s", "value": best, "unit": "s", "n_gpus": 1, "steps": 20, "warmup": max(args.warmup, 3),
            "ms_per_step": best * 1e3, "higher_is_better": False, "scaling": "weak", "vs_baseline": best / 0.23506,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": "benchmarks/jump_step_timing.py: 100 vmapped Heun+VBT SDE solves, ClipStepSizeController(PID, "
                                   "step_ts=129), SaveAt(ts=129), wall time of 3 runs, best of 20",
                       "accepted_steps_per_solve": acc, "published": "0.23506 s, hardware unstated (BASELINE.md §1)",
                       "prepared_3_runs_s": best_prepared, "device_ms_per_solve": device_ms},
            "gpu_launches": 3}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2")
    ap.add_argument("--scaling", default="strong", choices=["strong", "weak"],
                    help="strong: the workload's trajectory count is the total over all GPUs (BASELINE: 1M sharded across 1/2/4/8); "
                         "weak: it is the count per GPU")
    ap.add_argument("--trajectories", type=int, default=0, help="trajectory count (total if strong, per GPU if weak; default: the workload's)")
    ap.add_argument("--cpu-sample", type=int, default=0, help="trajectories of the cpu_baseline sample (default: per workload)")
    ap.add_argument("--ref-sample", type=int, default=0, help="trajectories per step of --impl reference (default: per workload)")
    ap.add_argument("--no-extras", action="store_true", help="headline only: no config_results (N=1) / weak field (N>1)")
    args = ap.parse_args()
    if args.workload == "published_jump_step":
        run_published_jump_step(args)
    elif args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
