#!/usr/bin/env python
"""Condense an .ncu-rep (via `ncu -i rep --page raw --csv`) into the handful of numbers we track per round.
usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep [kernel-index]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
idx = int(sys.argv[2]) if len(sys.argv) > 2 else 0
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, vals = rows[0], rows[1], rows[2 + idx]
d = {h: (v, u) for h, u, v in zip(hdr, units, vals)}


def num(k):
    v = d.get(k, ("nan", ""))[0].replace(",", "")
    try:
        return float(v)
    except ValueError:
        return float("nan")


print("kernel:", d.get("Kernel Name", ("?",))[0][:100], "grid", d.get("Grid Size", ("?",))[0], "block", d.get("Block Size", ("?",))[0])
keys = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "launch__registers_per_thread", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "sm__cycles_elapsed.max", "smsp__warps_eligible.avg.per_cycle_active",
        "smsp__warps_active.avg.per_cycle_active", "sm__inst_executed.avg.per_cycle_elapsed"]
for k in keys:
    if k in d:
        print(f"{k:75s} {d[k][0]:>18s} {d[k][1]}")
st = [(num(h), h) for h in d if h.startswith("smsp__pcsamp_warps_issue_stalled_") and not h.endswith("_not_issued")]
tot = sum(x for x, _ in st if x == x)
print("stall samples:")
for x, h in sorted(st, reverse=True)[:10]:
    print(f"  {h[len('smsp__pcsamp_warps_issue_stalled_'):]:28s} {x:10.0f} {100 * x / tot:5.1f}%")
