#!/usr/bin/env python
"""bench.py - headline benchmark of the ensemble integrator (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--workload c2|c1|c3|c5_heun|c5_shark]

metric   accepted RK steps / s across the ensemble (whole job, all ranks)
step     ONE pass of the hot path over one batch of synthetic input: a complete
         `diffeqsolve` of the whole ensemble resident in HBM
workload C2 (BASELINE.json configs[1]): Lorenz sigma=10 rho=28 beta=8/3, Dopri5,
         PIDController(rtol=atol=1e-8), dt0=None, t in [0,2], SaveAt(t1=True), fp64,
         2^20 trajectories PER GPU (weak scaling: the path shards by trajectory, no data-path
         collective; NCCL only gathers final states / reduces statistics after the kernel)
e2e      same metric through the public `diffrax_b200.diffeqsolve` call with HOST (pinned)
         buffers: H2D of y0 and D2H of ys/ts/stats/result inside the timed region
roofline FP64 FMA pipe: achieved = attempted steps x 316 flop / device time of
         the ensemble kernel (CUDA events on the launch stream); peak = DFMA-chain
         microbenchmark measured live (MEASURED_PEAKS.json has no FP64 figure)
--impl reference  times the CPU restatement (oracle/, "port": the reference is pure Python on
         JAX and jax is not installable here) on all host cores, bounded sample per step.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FLOP_PER_ATTEMPTED_STEP = {"c2": 316, "c1": 228, "c3": 1536}
DEFAULT_TRAJECTORIES = {"c1": 1024, "c2": 1 << 20, "c3": 1 << 18, "c4": 1 << 16, "c5_heun": 1 << 20, "c5_shark": 1 << 20}


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
    if name == "c3":
        # Arenstorf initial condition perturbed by 1e-4 N(0,1) (1e-3 sends a third of the ensemble through lunar
        # near-collisions needing > 8192 steps); one period; dense output with max_steps = 768 (103 GB of dense buffers).
        rng = np.random.default_rng(2 + seed_offset)
        y0 = np.array([0.994, 0.0, 0.0, -2.00158510637908252]) + 1e-4 * rng.standard_normal((n, 4))
        return dict(base, field="cr3bp", params=[0.012277471], solver="dopri8", y0=y0, t0=0.0, t1=17.0652165601579625,
                    rtol=1e-12, atol=1e-12, save_dense=True, max_steps=768,
                    label="C3 CR3BP(Arenstorf+1e-4 N(0,1))/Dopri8/PID(1e-12,1e-12)/fp64/one period/SaveAt(dense), max_steps=768")
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
    summary (profiles/r01_<workload>_*_ncu_full.txt), in bytes per launch; None when no capture is committed."""
    import glob
    import re
    for f in sorted(glob.glob(os.path.join(ROOT, "profiles", f"r01_{workload}_*ncu_full.txt"))):
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


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path.  Diffrax itself needs
    jax/equinox (absent, no network), so this is the oracle port on all host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    args.ref_sample = min(args.ref_sample, DEFAULT_TRAJECTORIES[args.workload])
    w = workload(args.workload, args.ref_sample)
    for _ in range(args.warmup):
        cpu_port_rate(w, min(args.ref_sample, 8192))
    rates, times = [], []
    cores = 1
    for _ in range(args.steps):
        r, dt, cores = cpu_port_rate(w, args.ref_sample)
        rates.append(r); times.append(dt)
    total_t = sum(times)
    value = float(np.sum(np.array(rates) * np.array(times)) / total_t)
    sample = f"{args.ref_sample} of the workload's trajectories per step (same seed/config), all host cores"
    line = {"impl": "reference", "metric": "accepted_rk_steps_per_s", "value": value, "unit": "steps/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total_t / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64" if w["dtype"] == np.float64 else "f32", "data": "synthetic",
            "config": {"workload": w["label"], "trajectories_per_step": args.ref_sample},
            "cpu_baseline": {"value": value, "unit": "steps/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def run_ours(args):
    import torch
    import torch.distributed as dist
    from diffrax_b200 import _dist, _lib

    rank, local, world = _dist.init_from_env()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU path)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    L = _lib.lib()

    n_local = args.trajectories or DEFAULT_TRAJECTORIES[args.workload]
    w = workload(args.workload, n_local, seed_offset=1000 * rank)
    dfx, term, solver, ctrl, saveat = _ours_objects(w, dev)
    y0_dev = torch.tensor(w["y0"], device=dev)
    y0_host = torch.tensor(w["y0"]).pin_memory()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2

    # prepared call: descriptor + output buffers built once; each step is ONE C-ABI call
    plan = dfx.prepare(term, solver, w["t0"], w["t1"], w["dt0"], y0_dev, saveat=saveat, stepsize_controller=ctrl,
                       max_steps=w["max_steps"])
    def _tensors(x):
        if isinstance(x, torch.Tensor):
            yield x
        elif isinstance(x, dict):
            for v in x.values():
                yield from _tensors(v)
    out_bytes = sum(int(t.numel() * t.element_size()) for k in plan._keep for t in _tensors(k))
    do_e2e = out_bytes < (2 << 30)      # C3's 69 GB of dense output is not staged through the host
    _, term_h, _, _, _ = _ours_objects(w, None)

    def step_device():
        return plan(throw=False)

    def step_host():
        return dfx.diffeqsolve(term_h, solver, w["t0"], w["t1"], w["dt0"], y0_host, saveat=saveat,
                               stepsize_controller=ctrl, throw=False, device=local, max_steps=w["max_steps"])

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up (also pays cudaMallocAsync pool growth, module load) ----
    for _ in range(max(args.warmup, 3)):
        sol = step_device()
    torch.cuda.synchronize()
    acc_local = int(sol.stats["num_accepted_steps"].sum())
    att_local = int(sol.stats["num_steps"].sum())
    failed_local = int((sol.result != 0).sum())

    # ---- timed: K steps, per-step CUDA events on the launch stream, L2 flushed between steps ----
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    L.dfx_reset_launch_count()
    barrier()
    wall0 = time.perf_counter()
    for e0, e1 in ev:
        flush.fill_(1)  # L2 flush, outside the event pair
        e0.record()
        sol = step_device()
        e1.record()
    barrier()
    wall = time.perf_counter() - wall0
    launches = int(L.dfx_launch_count())
    clocks = sampler.stop() if rank == 0 else None
    dev_ms = sum(e0.elapsed_time(e1) for e0, e1 in ev)  # device time of the K solves

    # ---- e2e: same K steps through the host-buffer call (H2D + D2H inside the timed region) ----
    e2e_s, h2d, d2h = float("nan"), 0, 0
    if do_e2e:
        for _ in range(2):
            sh = step_host()
        barrier()
        per_step = []
        t_all = time.perf_counter()
        for _ in range(args.steps):
            t0 = time.perf_counter()
            sh = step_host()
            _ = int(sh.result[0])  # read the step's result on the host
            per_step.append(time.perf_counter() - t0)
        barrier()
        e2e_mean_s = (time.perf_counter() - t_all) / args.steps
        # K host-timed steps; the per-step MEDIAN is reported (one scheduler hiccup of the host process - seen once as a
        # single 180 ms step among 30 - would otherwise decide the figure); the plain mean is kept next to it
        e2e_s = float(np.median(per_step)) * args.steps
        h2d = y0_host.numel() * y0_host.element_size() + (0 if w["keys"] is None else w["keys"].nbytes)
        d2h = sum(int(t.numel() * t.element_size()) for t in (sh.ts, sh.ys, sh.result)) \
            + 3 * int(sh.stats["num_steps"].numel()) * 4

    # ---- max over ranks, totals over ranks ----
    t_max = torch.tensor([dev_ms, (e2e_s if do_e2e else 0.0) * 1e3, wall * 1e3], dtype=torch.float64, device=dev)
    tot = torch.tensor([acc_local, att_local, failed_local], dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(t_max, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
        # the only communication of the path: gather final states + reduce statistics (not timed)
        _ = _dist.gather_final_states(sol.y_final, n_local * world)
    dev_ms_max, e2e_ms_max, wall_ms_max = (float(x) for x in t_max)
    acc, att, failed = (int(x) for x in tot)

    if rank == 0:
        ms_per_step = dev_ms_max / args.steps
        value = acc / (ms_per_step * 1e-3)
        flop = FLOP_PER_ATTEMPTED_STEP.get(args.workload)
        if flop is not None and args.workload != "c4":
            peak = float(L.dfx_measure_fma_peak(_lib.F64, local))  # TFLOP/s, live DFMA-chain microbenchmark
            achieved = (att / world) * flop / (ms_per_step * 1e-3) / 1e12  # per-GPU: the kernel of ONE rank
            roof = {"bound": "fma_fp64", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                    "frac": achieved / peak if peak > 0 else None, "traffic": None,
                    "peak_source": "dfx_measure_fma_peak: 8 independent DFMA chains/thread, 8 CTAs x 256 thr/SM, "
                                   "measured in this run (MEASURED_PEAKS.json holds no FP64 FMA figure)",
                    "flop_per_attempted_step": flop}
            if args.workload == "c3":   # dense output: 8 (16 d + 1) = 520 B per accepted step + inf padding of the tails
                by = float(out_bytes)
                roof["hbm_write"] = {"bytes_per_launch": by, "achieved_GBps": by / (ms_per_step * 1e-3) / 1e9 / world,
                                     "peak_GBps": 6546.6, "peak_source": "MEASURED_PEAKS.json hbm_gbs (copy, read+write)"}
        elif args.workload == "c4":
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
            peak = 2 * 8192.0 ** 3 / (best * 1e-3) / 1e12
            achieved = (att / world) * flop / (ms_per_step * 1e-3) / 1e12
            issued = (att / world) * 7 * 3 * 2 * 128 * 128 / (ms_per_step * 1e-3) / 1e12
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
            # ALGORITHM executes, per step, only the levels below the common prefix.  Counted exactly for this time grid:
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
            per_level, leaf = {"c5_heun": (3, 1), "c5_shark": (7, 4)}[args.workload]
            blocks = per_level * lv_total / nsteps_grid + leaf
            blocks_nocache = {"c5_heun": 3 * 8 + 4, "c5_shark": 7 * 8 + 9}[args.workload]
            ops = blocks * 77.0          # 20 x (add, rotate, xor) + 17 injection adds per block
            peak = float(L.dfx_measure_int_peak(local))
            achieved = (att / world) * ops / (ms_per_step * 1e-3) / 1e12
            roof = {"bound": "int32_alu", "achieved": achieved, "peak": peak, "unit": "Tera-op/s",
                    "frac": achieved / peak if peak > 0 else None, "traffic": None,
                    "peak_source": "dfx_measure_int_peak: add/rotate/xor chains, measured in this run",
                    "threefry_blocks_per_step": blocks, "threefry_blocks_per_step_without_descent_cache": blocks_nocache}
        roof["traffic"] = ncu_traffic(args.workload)   # (a capture of this very workload, or none)
        cpu_sample = min(args.cpu_sample, n_local)
        cpu_rate, cpu_dt, cores = cpu_port_rate(w, cpu_sample)
        line = {
            "metric": "accepted_rk_steps_per_s", "value": value, "unit": "steps/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64" if w["dtype"] == np.float64 else "f32", "data": "synthetic",
            "config": {"workload": w["label"], "trajectories_per_gpu": n_local, "trajectories_total": n_local * world,
                       "accepted_steps_per_solve": acc, "attempted_steps_per_solve": att, "failed_trajectories": failed,
                       "l2": "flushed between timed steps (256 MiB write outside the per-step CUDA-event pairs)",
                       "parallelism": f"trajectory-sharded x{world}, no data-path collective"},
            "e2e": ({"value": acc / (e2e_ms_max * 1e-3 / args.steps), "unit": "steps/s", "h2d_bytes_per_step": h2d,
                     "d2h_bytes_per_step": d2h, "ms_per_step": e2e_ms_max / args.steps, "statistic": "median of the K per-step wall times",
                     "mean_ms_per_step": e2e_mean_s * 1e3} if do_e2e else
                    {"value": None, "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0,
                     "note": "outputs exceed 2 GiB; not staged through the host"}),
            "gpu_launches": launches,
            "roofline": roof,
            "cpu_baseline": {"value": cpu_rate, "unit": "steps/s", "cores": cores, "kind": "port",
                             "sample": f"first {cpu_sample} trajectories of rank 0's batch, oracle (C port), "
                                       f"{cpu_dt:.2f} s"},
            "clocks": clocks, "wall_ms_per_step_incl_flush_and_host": wall_ms_max / args.steps,
        }
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
    ap.add_argument("--trajectories", type=int, default=0, help="trajectories per GPU (default: the workload's size)")
    ap.add_argument("--cpu-sample", type=int, default=1 << 20, help="trajectories of the cpu_baseline sample")
    ap.add_argument("--ref-sample", type=int, default=1 << 19, help="trajectories per step of --impl reference")
    args = ap.parse_args()
    if args.workload == "published_jump_step":
        run_published_jump_step(args)
    elif args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
