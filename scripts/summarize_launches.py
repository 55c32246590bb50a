"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: the launches of the LAST step, per kernel.
    python scripts/summarize_launches.py gpurun_out/launches.csv <launches_per_step>
"""
import csv
import re
import sys
from collections import OrderedDict

path, per_step = sys.argv[1], int(sys.argv[2])
rows = []
with open(path, newline="") as f:
    lines = [l for l in f if not l.startswith("==")]
for r in csv.DictReader(lines):
    if r.get("Metric Name") == "gpu__time_duration.sum":
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        us = v / 1e3 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1e3)
        name = re.sub(r"\(.*", "", r["Kernel Name"])
        rows.append((name, us))
rows = rows[-per_step:]
agg = OrderedDict()
for n, us in rows:
    a = agg.setdefault(n, [0, 0.0])
    a[0] += 1
    a[1] += us
tot = sum(a[1] for a in agg.values())
print(f"step total under ncu: {tot / 1e3:.3f} ms over {len(rows)} launches\n")
print("| kernel | launches | sum us | avg us | share |\n|---|---:|---:|---:|---:|")
for n, (c, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"| {n} | {c} | {us:.1f} | {us / c:.2f} | {100 * us / tot:.1f}% |")
