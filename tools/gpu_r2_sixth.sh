#!/bin/bash
mkdir -p gpurun_out
echo "== suite"
timeout -k 10 1500 python -m pytest tests -m gpu -q -p no:cacheprovider --maxfail=40 > gpurun_out/suite6.log 2>&1
echo "suite exit $?"; tail -n 3 gpurun_out/suite6.log; grep -E "^(FAILED|ERROR)" gpurun_out/suite6.log | head -40
echo "== bench"
timeout 1200 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_v5.json 2> gpurun_out/bench_v5.err; echo "bench exit $?"; tail -n 3 gpurun_out/bench_v5.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_v5.json"))
print("ms/step", d["ms_per_step"], "value", d["value"], "e2e", d["e2e"]["value"], "conv", d["roofline"]["achieved"], d["roofline"]["frac"], "share", d["roofline"]["share_of_step"], "clock", d["clocks"])
for s in d["roofline"]["secondary"]: print(s["kernel"][:30], round(s["achieved"], 1), round(s["frac"], 3), round(s["share_of_step"], 4))
print("cpu", d["cpu_baseline"])
for k, v in d["other_configs"].items(): print(k, round(v["value"], 3), v["unit"], round(v["ms_per_call"], 2), "ms", v.get("ms_per_call_eager"), round(v["algorithmic_tflops"], 1), "TF/s", round(v["frac_of_tensor_peak"], 3), "e2e", round(v["e2e"]["value"], 3), v.get("cpu_baseline", {}).get("value"))
PY
echo "== reference arm"
timeout 900 python bench.py --impl reference --steps 10 --warmup 3 > gpurun_out/bench_ref_v5.json 2>/dev/null; cut -c1-400 gpurun_out/bench_ref_v5.json
echo "== smoke"
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 2
