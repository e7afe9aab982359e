#!/bin/bash
# Lean epilogue with the accumulator read of the next chunk issued ahead (B200_TMEM_PIPELINE 1/0).
mkdir -p gpurun_out
timeout -k 10 600 python -m pytest tests/test_kernels_gpu.py -q -p no:cacheprovider -x 2>&1 | tail -n 2
for cfg in "B200_TMEM_PIPELINE=1" "B200_TMEM_PIPELINE=0"; do
  echo "== $cfg"
  env $cfg timeout 600 python tools/gemm_probe.py 2>&1 | tail -n 15 | head -n 11
  env $cfg timeout 900 python tools/splitk_ab.py _SPLIT_K 2>&1 | tail -n 2
done
