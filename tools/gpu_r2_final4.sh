#!/bin/bash
# Final round-2 evidence at HEAD: the -m gpu suite, both bench arms as the driver runs them, smoke, launch lists.
mkdir -p gpurun_out
timeout -k 10 1500 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/final4_gpu_tests.log 2>&1; echo "gpu tests exit $?"; tail -n 3 gpurun_out/final4_gpu_tests.log
timeout -k 10 900 python bench.py --impl reference --steps 10 --warmup 3 > gpurun_out/final4_bench_ref.json 2>/dev/null; cut -c1-300 gpurun_out/final4_bench_ref.json; echo
timeout -k 10 1500 python bench.py --steps 20 --warmup 5 > gpurun_out/final4_bench.json 2> gpurun_out/final4_bench.err; echo "bench exit $?"
python - <<'PY'
import json
d = json.load(open("gpurun_out/final4_bench.json"))
print("ms/step", d["ms_per_step"], "value", d["value"], "e2e", d["e2e"]["value"], "conv", d["roofline"]["achieved"], d["roofline"]["frac"], "share", d["roofline"]["share_of_step"], "clock", d["clocks"], "launches", d["gpu_launches"])
for s in d["roofline"]["secondary"]: print(s["kernel"][:30], round(s["achieved"], 1), round(s["frac"], 3), round(s["share_of_step"], 4))
print("cpu", d["cpu_baseline"])
for k, v in d["other_configs"].items(): print(k, round(v["value"], 3), v["unit"], round(v["ms_per_call"], 2), "ms", v.get("ms_per_call_eager"), round(v["algorithmic_tflops"], 1), "TF/s", round(v["frac_of_tensor_peak"], 3), "e2e", round(v["e2e"]["value"], 3), v.get("cpu_baseline", {}).get("value"))
PY
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final4_smoke.log 2>&1; tail -n 1 gpurun_out/final4_smoke.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1000 --csv --log-file gpurun_out/launches_smoke_final4.csv \
  python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/ncu_smoke2.log 2>&1; python tools/summarize_launches.py gpurun_out/launches_smoke_final4.csv | head -24
timeout -k 10 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_bench_final4.csv \
  python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-other-configs > gpurun_out/ncu_bench2.log 2>&1; python tools/summarize_launches.py gpurun_out/launches_bench_final4.csv | head -24
