#!/bin/bash
# Re-entry check at HEAD (one-launch split-K had not been through the whole suite): the -m gpu suite, the split-K A/B,
# and ncu launch lists of one C2 / C5 forward for the latency-bound work.
mkdir -p gpurun_out
timeout -k 10 1500 python -m pytest tests -m gpu -q -p no:cacheprovider --durations=15 > gpurun_out/suite_ninth.log 2>&1; echo "suite exit $?"; tail -n 25 gpurun_out/suite_ninth.log; grep -E "^(FAILED|ERROR)" gpurun_out/suite_ninth.log | head
timeout 600 python tools/splitk_ab.py _SPLIT_FUSED 2>&1 | tail -n 4
for w in c2 c5; do
  timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/launches_$w.csv python tools/one_forward.py $w > gpurun_out/launches_$w.log 2>&1
  echo "ncu $w exit $?"; python tools/summarize_launches.py gpurun_out/launches_$w.csv | head -24
done
