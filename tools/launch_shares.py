"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel: launches, total and share of
device time.  Usage: python tools/launch_shares.py file.csv [first_id last_id]"""
import csv
import re
import sys
from collections import OrderedDict

rows = []
with open(sys.argv[1]) as f:
    lines = [l for l in f if l.startswith('"')]
for r in csv.DictReader(lines):
    if r.get("Metric Name") == "gpu__time_duration.sum":
        rows.append((int(r["ID"]), r["Kernel Name"], float(r["Metric Value"].replace(",", "")), r["Grid Size"]))
if len(sys.argv) > 3:
    lo, hi = int(sys.argv[2]), int(sys.argv[3])
    rows = [r for r in rows if lo <= r[0] <= hi]
agg = OrderedDict()
for _, name, ns, grid in rows:
    short = re.sub(r"\(.*", "", name).replace("void ", "").replace("<unnamed>::", "")
    a = agg.setdefault(short, [0, 0.0])
    a[0] += 1
    a[1] += ns
tot = sum(a[1] for a in agg.values())
print(f"{len(rows)} launches, {tot / 1e6:.3f} ms of device time")
for k, (n, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{ns / tot * 100:6.2f} %  {n:5d} x  {ns / n / 1e3:10.1f} us  {k[:110]}")
