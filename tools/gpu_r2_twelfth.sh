#!/bin/bash
# ncu launch lists of one C5 / C2 (batch 32) / C2 (batch 1) UNet forward joined with the implicit-GEMM problem list:
# achieved TFLOP/s per problem class, to find what keeps the mid-size configurations at 0.2-0.35 of the tensor peak.
mkdir -p gpurun_out
for w in c5 c2n32 c2; do
  timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/launches_$w.csv python tools/one_forward.py $w > gpurun_out/launches_$w.log 2>&1
  echo "ncu $w exit $?"; python tools/join_shapes.py gpurun_out/launches_$w.csv gpurun_out/shapes_$w.json
done
