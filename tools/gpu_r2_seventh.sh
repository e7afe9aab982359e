#!/bin/bash
mkdir -p gpurun_out
echo "== suite (relaxed pair threshold)"
timeout -k 10 1500 python -m pytest tests -m gpu -q -p no:cacheprovider --maxfail=40 > gpurun_out/suite7.log 2>&1
echo "suite exit $?"; tail -n 3 gpurun_out/suite7.log; grep -E "^(FAILED|ERROR)" gpurun_out/suite7.log | head -40
echo "== ncu conv pair"
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:igemm_tc_kernel -s 2 -c 1 -f \
  -o gpurun_out/prof_igemm_pair python tools/conv_probe.py > gpurun_out/ncu_igemm_pair.log 2>&1; echo "ncu conv exit $?"; tail -n 2 gpurun_out/ncu_igemm_pair.log
echo "== launch list of one bench step"
timeout -k 10 1500 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_bench_r2.csv \
  python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-other-configs > gpurun_out/ncu_bench_r2.log 2>&1; echo "ncu bench exit $?"
python tools/summarize_launches.py gpurun_out/launches_bench_r2.csv | head -30
echo "== other configs quick (C2/C5 after threshold change)"
timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_v6.json 2>/dev/null
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_v6.json"))
print("C3 ms/step", round(d["ms_per_step"], 1), {k: (round(v["ms_per_call"], 2), v.get("ms_per_call_eager")) for k, v in d["other_configs"].items()})
PY
