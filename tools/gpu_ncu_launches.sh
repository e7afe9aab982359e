#!/bin/bash
mkdir -p gpurun_out
timeout -k 10 1500 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_c3.csv \
  python tools/perf_c3.py --shape 160,224,160 --iters 0 --breakdown 0 > gpurun_out/ncu_c3.log 2>&1
tail -3 gpurun_out/ncu_c3.log
python - <<'PY'
import csv, collections, re
rows = []
with open('gpurun_out/launches_c3.csv') as f:
    lines = [l for l in f if not l.startswith('==')]
r = csv.DictReader(lines)
agg = collections.defaultdict(lambda: [0, 0.0])
for row in r:
    name = row.get('Kernel Name', '')
    try:
        v = float(row['Metric Value'].replace(',', ''))
    except Exception:
        continue
    unit = row.get('Metric Unit', 'ns')
    ms = v / 1e6 if unit in ('ns', 'nsecond') else v / 1e3 if unit in ('us', 'usecond') else v
    short = re.sub(r'\(.*', '', name)[:70]
    agg[short][0] += 1
    agg[short][1] += ms
tot = sum(v[1] for v in agg.values())
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:25]:
    print(f"{v[1]:10.2f} ms  x{v[0]:5d}  {100*v[1]/tot:5.1f}%  {k}")
print(f"total {tot:.1f} ms")
PY
