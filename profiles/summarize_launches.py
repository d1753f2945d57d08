"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name.

    python profiles/summarize_launches.py gpurun_out/launches.csv [--last-step N]  > profiles/r01_launches_summary.md
"""
import csv
import re
import sys
from collections import defaultdict

path = sys.argv[1]
rows = []
with open(path, newline="") as f:
    lines = [l for l in f if not l.startswith("==")]
for r in csv.DictReader(lines):
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    try:
        v = float(r["Metric Value"].replace(",", ""))
    except ValueError:
        continue
    unit = r.get("Metric Unit", "ns")
    us = v / 1e3 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1e3)
    name = re.sub(r"\(.*", "", r["Kernel Name"])
    name = re.sub(r"^void ", "", name)
    rows.append((int(r["ID"]), name, us))
agg = defaultdict(lambda: [0, 0.0])
for _, n, us in rows:
    agg[n][0] += 1
    agg[n][1] += us
total = sum(v[1] for v in agg.values())
print(f"launches: {len(rows)}, total kernel time {total / 1e3:.2f} ms (cold-cache, serialised by ncu: compare SHARES)\n")
print("| kernel | launches | total ms | share | mean us |")
print("|---|---:|---:|---:|---:|")
for n, (c, us) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
    print(f"| `{n[:90]}` | {c} | {us / 1e3:.3f} | {100 * us / total:.1f}% | {us / c:.1f} |")
