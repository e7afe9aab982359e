"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count, total and share."""
import csv, re, sys
from collections import defaultdict
rows = []
with open(sys.argv[1]) as f:
    lines = [l for l in f if not l.startswith("==")]
rd = csv.DictReader(lines)
tot = defaultdict(lambda: [0, 0.0])
for r in rd:
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = re.sub(r"\(.*", "", r["Kernel Name"])
    v = float(r["Metric Value"].replace(",", ""))
    unit = r.get("Metric Unit", "ns")
    us = v / 1e3 if unit in ("ns", "nsecond") else v if unit in ("us", "usecond") else v * 1e3
    tot[name][0] += 1
    tot[name][1] += us
total = sum(v[1] for v in tot.values())
n = sum(v[0] for v in tot.values())
print(f"{n} kernels, {total/1e3:.3f} ms (serialised, cold-cache: compare shares)")
for name, (c, us) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print(f"{us:10.1f} us {100*us/total:5.1f}%  x{c:<4d} avg {us/c:7.1f} us  {name[:90]}")
