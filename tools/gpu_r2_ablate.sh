#!/bin/bash
# Attention ablation matrix at T = S = 89 600, d = 512: {single-CTA replay, CTA-pair} x {full, 7 ablations}.
mkdir -p gpurun_out
: > gpurun_out/attn_ablate.log
for pair in 0 1; do
  for ab in base 1 3 4; do
    if [ $ab = base ]; then lib=""; else lib="tools/devlibs/libb200gen_ab$ab.so"; fi
    res=$(timeout -k 10 120 env B200_FLASH_PAIR=$pair B200_DEV_LIB=$lib python tools/attn_probe.py replay 2>&1 | grep "ms," | awk '{print $6}' | tr '\n' ' ')
    echo "pair=$pair ablate=$ab : $res" | tee -a gpurun_out/attn_ablate.log
  done
done
echo "== attention tests"
timeout -k 10 300 python -m pytest tests/test_kernels_gpu.py tests/test_fullsize_gpu.py -q -p no:cacheprovider -k "attention" 2>&1 | tail -n 3
timeout -k 10 300 env B200_FLASH_PAIR=1 python -m pytest tests/test_kernels_gpu.py tests/test_fullsize_gpu.py -q -p no:cacheprovider -k "attention" 2>&1 | tail -n 3
