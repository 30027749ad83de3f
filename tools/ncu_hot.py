#!/usr/bin/env python
"""Where a kernel's executed instructions and stall samples go, by CUDA source line (needs -lineinfo and --import-source on):
    python tools/ncu_hot.py report.ncu-rep [top-n]"""
import collections
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
ex_by_line, smp_by_line, mix = collections.Counter(), collections.Counter(), collections.Counter()
tot_ex = tot_smp = 0
fname, h = "?", None
for r in rows:
    if len(r) >= 2 and r[0] == "File Path":
        fname = r[1].split("/")[-1]
        continue
    if r and r[0] == "Line No":
        h = r
        i_ex, i_smp = h.index("Instructions Executed"), h.index("# Samples")
        continue
    if h is None or len(r) <= i_ex:
        continue
    try:
        ex, smp = int(r[i_ex]), int(r[i_smp])
    except ValueError:
        continue
    if r[0].isdigit():                      # a CUDA source line: aggregated over its SASS
        key = f"{fname}:{r[0]}  {r[1].strip()[:120]}"
        ex_by_line[key] += ex
        smp_by_line[key] += smp
        tot_ex += ex
        tot_smp += smp
    elif r[2].startswith("0x"):             # one SASS instruction (the view elides most of them)
        w = r[3].split()
        if w:
            mix[w[1] if w[0].startswith("@") and len(w) > 1 else w[0]] += ex
print(f"warp instructions executed: {tot_ex}, samples: {tot_smp}")
print("top source lines by executed instructions:")
for k, v in ex_by_line.most_common(top):
    print(f"  {100 * v / max(tot_ex, 1):5.1f}%  smp {100 * smp_by_line[k] / max(tot_smp, 1):5.1f}%  {k}")
