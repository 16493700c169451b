#!/usr/bin/env python
"""Summarise an .ncu-rep (raw + source pages) into a small text file for profiles/."""
import collections, csv, re, subprocess, sys

rep = sys.argv[1]
raw = list(csv.reader(subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout.splitlines()))
hdr, units, vals = raw[0], raw[1], raw[2]
want = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tc.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__warps_eligible.avg.per_cycle_active", "sm__cycles_elapsed.avg", "smsp__cycles_active.avg",
        "local_load", "local_store"]
print(f"# {rep}")
for h, u, v in zip(hdr, units, vals):
    if h in want or h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio") or \
       (h.startswith("smsp__average_warp_latency_issue_stalled") and h.endswith(".ratio")):
        print(f"{h} [{u}] = {v}")
src = list(csv.reader(subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout.splitlines()))
h = src[1]
ia, ie, ist = h.index("Source"), h.index("Instructions Executed"), h.index("Warp Stall Sampling (All Samples)")
ops, stalls, total = collections.Counter(), collections.Counter(), 0
for r in src[2:]:
    try:
        n, st = int(r[ie]), int(r[ist])
    except Exception:
        continue
    m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[ia])
    op = (m.group(2) if m else r[ia][:10]).split(".")[0]
    ops[op] += n; stalls[op] += st; total += n
print(f"\n# SASS opcode mix (warp-level instructions executed, total {total})")
for k, v in ops.most_common(22):
    print(f"{k:10s} {v:14d} {100 * v / total:6.2f}%   stall samples {stalls[k]}")
fp64 = sum(v for k, v in ops.items() if k in ("DFMA", "DADD", "DMUL", "DSETP", "DMNMX", "F2F"))
print(f"FP64-pipe instructions: {fp64} ({100 * fp64 / total:.1f}% of all)")
