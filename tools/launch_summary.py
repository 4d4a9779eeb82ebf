"""Summarise an ncu launch list (`ncu --metrics gpu__time_duration.sum --csv --log-file X.csv`): per kernel name the
launch count, total and mean duration (us) and share of the sum.  usage: python tools/launch_summary.py X.csv [top]"""
import csv
import re
import sys
from collections import defaultdict

rows = []
with open(sys.argv[1]) as f:
    lines = [l for l in f if not l.startswith("==")]
rd = csv.DictReader(lines)
tot = defaultdict(lambda: [0, 0.0])
order = []
for r in rd:
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = re.sub(r"\(.*", "", r["Kernel Name"])
    v = float(r["Metric Value"].replace(",", ""))
    unit = r["Metric Unit"]
    us = v / 1e3 if unit in ("ns", "nsecond") else v * 1e3 if unit in ("ms", "msecond") else v
    tot[name][0] += 1
    tot[name][1] += us
    order.append((name, us))
s = sum(v[1] for v in tot.values())
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
print(f"{len(order)} launches, {s:.1f} us in total")
for name, (c, t) in sorted(tot.items(), key=lambda kv: -kv[1][1])[:top]:
    print(f"{t:10.1f} us {100 * t / s:5.1f} %  x{c:<4d} mean {t / c:8.1f} us  {name}")
