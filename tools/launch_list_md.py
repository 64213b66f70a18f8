#!/usr/bin/env python
"""Turn an `ncu --metrics gpu__time_duration.sum --csv` launch list into the per-kernel table kept under profiles/.

    python tools/launch_list_md.py gpurun_out/r2_launches.csv "<command that was profiled>" > profiles/r2_launches.md
"""
import csv
import sys
from collections import OrderedDict

path, cmd = sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else "python bench.py"
rows = [r for r in csv.reader(open(path)) if len(r) > 10]
head = rows[0]
ik, iv, ip, iu = head.index("Kernel Name"), head.index("Metric Value"), head.index("Process Name"), head.index("Metric Unit")
agg = OrderedDict()
for r in rows[1:]:
    v = float(r[iv].replace(",", ""))
    v = v / 1e3 if r[iu] in ("ns", "nsecond") else v * 1e3 if r[iu] in ("ms", "msecond") else v
    key = (r[ip], r[ik])
    a = agg.setdefault(key, [0, 0.0])
    a[0] += 1
    a[1] += v
total = sum(a[1] for a in agg.values())
print(f"# ncu launch list of `{cmd}`\n")
print("`ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv` — per-launch times are cold-cache and")
print(f"serialised; only each kernel's SHARE is meaningful.  Raw CSV next to this file.  {sum(a[0] for a in agg.values())} launches, {total / 1e3:.1f} ms in total.\n")
print("| process | kernel | launches | total us | mean us | share of GPU time |")
print("|---|---|---|---|---|---|")
for (proc, k), (n, t) in agg.items():
    print(f"| {proc} | `{k[:110]}` | {n} | {t:.1f} | {t / n:.1f} | {100 * t / total:.1f} % |")
