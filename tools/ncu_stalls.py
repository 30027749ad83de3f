#!/usr/bin/env python
"""Summarise an .ncu-rep (first kernel): headline metrics, stall reasons per issue, and where no-instruction stalls fall
relative to 128 B / 256 B instruction-fetch boundaries.   python tools/ncu_stalls.py report.ncu-rep"""
import collections
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
h, r = rows[0], rows[2]
for k in ["Kernel Name", "gpu__time_duration.sum", "launch__registers_per_thread", "smsp__issue_active.avg.pct_of_peak_sustained_active",
          "smsp__inst_executed.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__warps_active.avg.pct_of_peak_sustained_active"]:
    if k in h:
        print(f"{k:60s} {r[h.index(k)]} {rows[1][h.index(k)]}")
st = []
for i, x in enumerate(h):
    if "issue_stalled" in x and x.endswith("per_issue_active.ratio"):
        st.append((float(r[i]), x.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")))
print("stall cycles per issued instruction:")
for v, x in sorted(st, reverse=True)[:9]:
    print(f"  {x:28s} {v:7.3f}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
h, data = rows[1], rows[2:]
i_n, i_s = h.index("stall_no_inst"), h.index("# Samples")
print("SASS instructions:", len(data), f"({len(data) * 16 / 1024:.0f} KB)", " samples:", sum(int(d[i_s]) for d in data), " no_inst samples:", sum(int(d[i_n]) for d in data))
pos = collections.Counter()
for d in data:
    pos[(int(d[0], 16) // 16) % 16] += int(d[i_n])
print("no_inst samples by instruction slot within 256 B:", [pos[k] for k in range(16)])
