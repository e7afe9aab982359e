#!/bin/bash
# Per-problem-class view of ONE C3 forward at full size (where do the ~29 ms outside conv / attention / GroupNorm go?),
# and of the C5 / C2-batch-32 forwards after the second-half changes.
mkdir -p gpurun_out
for w in c3 c5 c2n32; do
  timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/launches_$w.csv python tools/one_forward.py $w > gpurun_out/launches_$w.log 2>&1
  echo "ncu $w exit $?"; python tools/join_shapes.py gpurun_out/launches_$w.csv gpurun_out/shapes_$w.json | head -48
done
