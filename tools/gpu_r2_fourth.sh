#!/bin/bash
# Round-2 fourth GPU call: programmatic dependent launch (suite + A/B on the latency-bound configs), launch lists of
# one C2 / C5 UNet forward, CUDA-event + ncu evidence for the HBM-bound kernels.
mkdir -p gpurun_out
echo "== suite with PDL"
timeout -k 10 1500 python -m pytest tests -m gpu -q -p no:cacheprovider --maxfail=40 > gpurun_out/suite4.log 2>&1
echo "suite exit $?"; tail -n 3 gpurun_out/suite4.log; grep -E "^(FAILED|ERROR)" gpurun_out/suite4.log | head -40
echo "== PDL A/B (other configs through bench.py)"
for pdl in 0 1; do
  timeout 900 env B200_PDL=$pdl python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/ab_pdl$pdl.json 2> gpurun_out/ab_pdl$pdl.err; echo "pdl=$pdl exit $?"
  python - <<PY
import json
d = json.load(open("gpurun_out/ab_pdl$pdl.json"))
print("PDL=$pdl  C3 ms/step", round(d["ms_per_step"], 1), " | ", {k: round(v["ms_per_call"], 2) for k, v in d["other_configs"].items()})
PY
done
echo "== HBM-bound kernels"
timeout 600 python tools/hbm_probe.py > gpurun_out/hbm_probe.log 2>&1; echo "probe exit $?"; cat gpurun_out/hbm_probe.log
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
  -k regex:"gn_apply_kernel|gn_partial_kernel|ddim_step_kernel|vq_argmin_kernel" -c 6 -f -o gpurun_out/prof_hbm \
  python tools/hbm_probe.py > gpurun_out/ncu_hbm.log 2>&1; echo "ncu hbm exit $?"; tail -n 2 gpurun_out/ncu_hbm.log
echo "== launch lists"
for w in c2 c5; do
  timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/launches_$w.csv python tools/one_forward.py $w > gpurun_out/launches_$w.log 2>&1
  echo "ncu $w exit $?"; python tools/summarize_launches.py gpurun_out/launches_$w.csv | head -24
done
ls -la gpurun_out/*.ncu-rep
