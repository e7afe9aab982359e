#!/bin/bash
mkdir -p gpurun_out
for i in 1 2 3; do
  timeout -k 10 1500 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/stab_$i.log 2>&1; echo "run $i exit $?"; tail -n 2 gpurun_out/stab_$i.log; grep -E "^(FAILED|ERROR)" gpurun_out/stab_$i.log | head
done
timeout -k 10 900 env B200_ACT_DTYPE=bf16 python -m pytest tests -m gpu -q -p no:cacheprovider -k "not c2_reference and not fullsize" > gpurun_out/stab_bf16.log 2>&1; echo "bf16 exit $?"; tail -n 2 gpurun_out/stab_bf16.log; grep -E "^(FAILED|ERROR)" gpurun_out/stab_bf16.log | head
