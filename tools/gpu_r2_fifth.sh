#!/bin/bash
# Round-2 fifth GPU call: suite + bench with the CTA-pair attention default, the HBM kernels after their rework (CUDA events +
# one ncu --set full capture each), one ncu --set full capture of the pair attention kernel, split-K A/B on the latent UNets.
mkdir -p gpurun_out
echo "== suite"
timeout -k 10 1500 python -m pytest tests -m gpu -q -p no:cacheprovider --maxfail=40 > gpurun_out/suite5.log 2>&1
echo "suite exit $?"; tail -n 3 gpurun_out/suite5.log; grep -E "^(FAILED|ERROR)" gpurun_out/suite5.log | head -40
echo "== bench"
timeout 1200 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_v4.json 2> gpurun_out/bench_v4.err; echo "bench exit $?"; tail -n 3 gpurun_out/bench_v4.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_v4.json"))
print("ms/step", d["ms_per_step"], "value", d["value"], "e2e", d["e2e"]["value"], "conv", d["roofline"]["achieved"], d["roofline"]["frac"], "share", d["roofline"]["share_of_step"], "clock", d["clocks"])
for s in d["roofline"]["secondary"]: print(s["kernel"][:30], round(s["achieved"], 1), round(s["frac"], 3), round(s["share_of_step"], 4))
print("cpu", d["cpu_baseline"])
for k, v in d["other_configs"].items(): print(k, round(v["value"], 3), v["unit"], round(v["ms_per_call"], 2), "ms", round(v["algorithmic_tflops"], 1), "TF/s", round(v["frac_of_tensor_peak"], 3), "e2e", round(v["e2e"]["value"], 3), v.get("cpu_baseline", {}).get("value"))
PY
echo "== HBM-bound kernels (CUDA events)"
timeout 600 python tools/hbm_probe.py > gpurun_out/hbm_probe2.log 2>&1; echo "probe exit $?"; cat gpurun_out/hbm_probe2.log
for k in gn_partial_kernel gn_apply_kernel ddim_step_kernel vq_argmin; do
  timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:$k -s 1 -c 1 -f \
    -o gpurun_out/prof_$k python tools/hbm_probe.py > gpurun_out/ncu_$k.log 2>&1; echo "ncu $k exit $?"
done
echo "== ncu pair attention"
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:flash_pair -s 2 -c 1 -f \
  -o gpurun_out/prof_flash_pair python tools/attn_probe.py replay > gpurun_out/ncu_flash_pair.log 2>&1; echo "ncu flash exit $?"; tail -n 2 gpurun_out/ncu_flash_pair.log
echo "== split-K A/B"
timeout 600 python tools/splitk_ab.py _SPLIT_K > gpurun_out/splitk_ab2.log 2>&1; tail -n 3 gpurun_out/splitk_ab2.log
ls -la gpurun_out/*.ncu-rep
