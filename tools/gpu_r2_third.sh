#!/bin/bash
# Round-2 third GPU call: the CTA-pair attention kernel (correctness first, under a short timeout in case of a hang),
# its timing at T = S = 89 600, then the suite and the bench with the new defaults.
mkdir -p gpurun_out
echo "== attention tests (pair kernel)"
timeout -k 10 300 python -m pytest tests/test_kernels_gpu.py tests/test_fullsize_gpu.py -q -p no:cacheprovider -k "attention" > gpurun_out/attn_tests.log 2>&1
echo "attention tests exit $?"; tail -n 5 gpurun_out/attn_tests.log; grep -E "^(FAILED|ERROR)" gpurun_out/attn_tests.log | head
echo "== probe"
timeout -k 10 300 python tools/attn_probe.py > gpurun_out/attn_probe.log 2>&1; echo "probe exit $?"; tail -n 14 gpurun_out/attn_probe.log
echo "== suite"
timeout -k 10 1500 python -m pytest tests -m gpu -q -p no:cacheprovider --maxfail=40 > gpurun_out/suite3.log 2>&1
echo "suite exit $?"; tail -n 3 gpurun_out/suite3.log; grep -E "^(FAILED|ERROR)" gpurun_out/suite3.log | head -40
echo "== bench"
timeout 1200 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_v3.json 2> gpurun_out/bench_v3.err; echo "bench exit $?"; tail -n 3 gpurun_out/bench_v3.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_v3.json"))
print("ms/step", d["ms_per_step"], "value", d["value"], "e2e", d["e2e"]["value"], "conv", d["roofline"]["achieved"], d["roofline"]["frac"], "clock", d["clocks"])
for s in d["roofline"]["secondary"]: print(s["kernel"][:30], round(s["achieved"], 1), round(s["frac"], 3), round(s["share_of_step"], 4))
print("cpu", d["cpu_baseline"])
for k, v in d["other_configs"].items(): print(k, round(v["value"], 3), v["unit"], round(v["ms_per_call"], 2), "ms", round(v["algorithmic_tflops"], 1), "TF/s", round(v["frac_of_tensor_peak"], 3), "e2e", round(v["e2e"]["value"], 3), v.get("cpu_baseline", {}).get("value"))
PY
