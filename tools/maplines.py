"""Dev tool: per-source-line warp-stall samples of one kernel from an ncu report.

ncu's SASS source page (``ncu -i rep --page source --csv``) carries the sampling counters per instruction; ``nvdisasm -g``
of the same cubin (built with -lineinfo) carries the source line of every instruction.  Both list the kernel's
instructions in address order, so they are zipped and the samples summed per (file, line).

    cuobjdump -xelf all generativemodels_b200/lib/libb200gen.so          # -> igemm.sm_100a.cubin, ...
    nvdisasm -g -c igemm.sm_100a.cubin > igemm.sass
    ncu -i gpurun_out/prof.ncu-rep --page source --csv > prof_source.csv
    python tools/maplines.py igemm.sass _ZN4b20015igemm_tc_kernelILi256ELi6ELb1EEE prof_source.csv [top_n] [source_dir]
"""
import csv
import re
import sys
from pathlib import Path

sass, kern, rep = sys.argv[1], sys.argv[2], sys.argv[3]
top_n = int(sys.argv[4]) if len(sys.argv) > 4 else 30
src_dir = Path(sys.argv[5]) if len(sys.argv) > 5 else Path(__file__).resolve().parents[1] / "generativemodels_b200" / "csrc"
lines = open(sass).read().split("\n")
start = next(i for i, l in enumerate(lines) if l.startswith(".text." + kern))
cur, ins = None, []
for l in lines[start + 1:]:
    if l.startswith("//---------------------"):
        break
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m:
        ins.append((int(m.group(1), 16), cur))
rows = list(csv.reader(open(rep)))
h = rows[1]
ix = {n: i for i, n in enumerate(h)}
data = rows[2:]
print(rows[0][1][:90], f"| {len(ins)} instructions in the cubin, {len(data)} in the report")
base = int(data[0][ix["Address"]], 16)
stall_cols = [n for n in h if n.startswith("stall_") and "Not Issued" not in n]
byline = {}
for k, r in enumerate(data):
    off = int(r[ix["Address"]], 16) - base
    loc = ins[k][1] if k < len(ins) and ins[k][0] == off else None
    d = byline.setdefault(loc, [0, {}])
    d[0] += int(r[ix["# Samples"]] or 0)
    for n in stall_cols:
        v = int(r[ix[n]] or 0)
        if v:
            d[1][n] = d[1].get(n, 0) + v
tot = sum(v[0] for v in byline.values())
print("samples:", tot)
cache = {}
for loc, (smp, st) in sorted(byline.items(), key=lambda kv: -kv[1][0])[:top_n]:
    text = ""
    if loc and (src_dir / loc[0]).exists():
        if loc[0] not in cache:
            cache[loc[0]] = (src_dir / loc[0]).read_text().split("\n")
        if loc[1] - 1 < len(cache[loc[0]]):
            text = cache[loc[0]][loc[1] - 1].strip()[:95]
    top = sorted(st.items(), key=lambda kv: -kv[1])[:2]
    print(f"{smp:6d} {100 * smp / max(tot, 1):5.1f}%  {str(loc):28s} {text:95s} {top}")
