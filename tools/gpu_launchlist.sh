#!/bin/bash
mkdir -p gpurun_out
for w in brain c2; do
  timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/launches_$w.csv python tools/one_forward.py $w > gpurun_out/launches_$w.log 2>&1
  echo "ncu $w exit $?"; wc -l gpurun_out/launches_$w.csv
done
