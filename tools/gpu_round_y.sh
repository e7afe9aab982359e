#!/bin/bash
mkdir -p gpurun_out
timeout -k 10 600 python tools/decode_probe.py > gpurun_out/decode.log 2>&1; cat gpurun_out/decode.log | tail -8
timeout -k 10 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/decode_launches.csv -c 4000 python tools/decode_probe.py > /dev/null 2>&1
python - <<'PY'
import csv, collections, re
with open('gpurun_out/decode_launches.csv') as f:
    lines = [l for l in f if not l.startswith('==')]
agg = collections.defaultdict(lambda: [0, 0.0])
for row in csv.DictReader(lines):
    try: v = float(row['Metric Value'].replace(',', ''))
    except Exception: continue
    unit = row.get('Metric Unit', 'ns')
    us = v / 1e3 if unit in ('ns', 'nsecond') else v if unit in ('us', 'usecond') else v * 1e3
    k = re.sub(r'\(.*', '', row.get('Kernel Name', ''))[:60]
    agg[k][0] += 1; agg[k][1] += us
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:12]:
    print(f"{v[1]:10.0f} us  x{v[0]:5d}  avg {v[1]/v[0]:7.1f} us  {k}")
PY
