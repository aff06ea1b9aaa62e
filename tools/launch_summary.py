"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-launch table + per-kernel totals."""
import csv
import re
import sys
from collections import OrderedDict

path = sys.argv[1]
rows = []
with open(path) as f:
    lines = [ln for ln in f if not ln.startswith("==")]
for r in csv.DictReader(lines):
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = re.sub(r"\(.*", "", r["Kernel Name"])
    unit = r["Metric Unit"]
    v = float(r["Metric Value"].replace(",", ""))
    us = v / 1000.0 if unit in ("nsecond", "ns") else (v if unit in ("usecond", "us") else v * 1000.0)
    grid = r.get("Grid Size", "")
    rows.append((int(r["ID"]), name, grid, r.get("Block Size", ""), us))
total = sum(r[4] for r in rows)
verbose = "-v" in sys.argv
if verbose:
    for r in rows:
        print("%4d %-34s %-18s %-14s %9.1f us" % r)
agg = OrderedDict()
for _, name, _, _, us in rows:
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += us
print("total %.1f us over %d launches" % (total, len(rows)))
for name, (cnt, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("  %-40s x%-4d %9.1f us  %5.1f%%" % (name, cnt, us, 100 * us / total))
