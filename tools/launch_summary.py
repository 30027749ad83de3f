#!/usr/bin/env python
"""Summarise an ncu launch list (`ncu --metrics gpu__time_duration.sum --csv --log-file X.csv ...`) per kernel.
usage: python tools/launch_summary.py gpurun_out/launches.csv "command line that was profiled" """
import collections
import csv
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr = rows[0]
i_name, i_val, i_unit = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
tot = collections.Counter()
cnt = collections.Counter()
for r in rows[1:]:
    v = float(r[i_val].replace(",", ""))
    v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(r[i_unit], 1.0)
    tot[r[i_name]] += v
    cnt[r[i_name]] += 1
total = sum(tot.values())
print(f"# ncu launch list of `{sys.argv[2] if len(sys.argv) > 2 else '?'}` (gpu__time_duration.sum, cold-cache, serialised)")
print("# kernel | launches | total us | share")
for k, v in tot.most_common(12):
    print(f"{k[:90]} | {cnt[k]} | {v:.1f} | {100 * v / total:.1f}%")
