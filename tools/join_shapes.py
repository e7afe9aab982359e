"""Dev tool: join an ncu launch list (gpu__time_duration.sum per kernel, tools/gpu_launchlist.sh) with the implicit-GEMM
problem list tools/one_forward.py wrote for the same forward, and print achieved TFLOP/s per problem class.

    python tools/join_shapes.py gpurun_out/launches_c5.csv gpurun_out/shapes_c5.json
"""
import csv
import json
import sys
from collections import defaultdict

rows = list(csv.reader(open(sys.argv[1])))
hdr = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
ix = {n: i for i, n in enumerate(rows[hdr])}
launches = [(r[ix["Kernel Name"]], float(r[-1]) / 1e3, r[ix["Grid Size"]]) for r in rows[hdr + 2:] if len(r) > ix["Kernel Name"]]
shapes = json.load(open(sys.argv[2]))
ig = [(n, us, g) for (n, us, g) in launches if "igemm_tc_kernel" in n]
red = [us for (n, us, g) in launches if "split_reduce" in n]
if len(ig) != len(shapes):
    print(f"warning: {len(ig)} igemm launches vs {len(shapes)} logged problems")
agg = defaultdict(lambda: [0, 0.0, 0.0, ""])
ri = 0
for (n, us, g), s in zip(ig, shapes):
    extra = 0.0
    if s["split"] and ri < len(red):
        extra = red[ri]
        ri += 1
    key = (s["rows"], s["cout"], s["chunks"], s["split"])
    a = agg[key]
    a[0] += 1
    a[1] += us + extra
    a[2] += 2.0 * s["rows"] * s["cout"] * s["chunks"] * 64
    a[3] = n.split("igemm_tc_kernel")[1][:12] + " grid " + g
total = sum(a[1] for a in agg.values())
print(f"{len(ig)} implicit-GEMM launches, {total / 1e3:.3f} ms (+ reductions), all kernels {sum(u for _, u, _ in launches) / 1e3:.3f} ms")
for key, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"  {a[1]:9.1f} us  x{a[0]:<3d} rows={key[0]:<7d} cout={key[1]:<5d} K={key[2] * 64:<6d} {'split ' if key[3] else '      '}"
          f"{a[2] / a[1] / 1e6:8.1f} TFLOP/s (padded K)  {a[3]}")
other = defaultdict(lambda: [0, 0.0])
for (n, us, g) in launches:
    if "igemm_tc_kernel" not in n:
        k = n.split("(")[0][-40:]
        other[k][0] += 1
        other[k][1] += us
for k, (c, us) in sorted(other.items(), key=lambda kv: -kv[1][1])[:14]:
    print(f"  {us:9.1f} us  x{c:<3d} {k}")
