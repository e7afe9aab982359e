#!/bin/bash
# Split-K after the reduction kernel got eight loads in flight: per-shape A/B and UNet steps for three minimum range
# lengths (B200_SPLIT_RANGE_MIN chunks per range), the K=256 narrowing default, layernorm (128-bit), and a fresh
# source-level capture of the feed-forward GEMM with the lean epilogue.
mkdir -p gpurun_out
timeout -k 10 600 python -m pytest tests/test_kernels_gpu.py -q -p no:cacheprovider -x 2>&1 | tail -n 2
for r in 4 12 24; do
  echo "== B200_SPLIT_RANGE_MIN=$r"
  B200_SPLIT_RANGE_MIN=$r timeout 600 python tools/gemm_probe.py --split-ab 2>&1 | tail -n 15
  B200_SPLIT_RANGE_MIN=$r timeout 900 python tools/splitk_ab.py _SPLIT_K 2>&1 | tail -n 4
done
echo "== fused, B200_SPLIT_RANGE_MIN=12"; B200_SPLIT_RANGE_MIN=12 timeout 900 python tools/splitk_ab.py _SPLIT_FUSED 2>&1 | tail -n 4
timeout -k 10 600 ncu --profile-from-start off --set full --clock-control none --import-source on -f -o gpurun_out/prof_gemm_ff2 \
  python tools/gemm_probe.py --profile 32768,2048,256 > gpurun_out/ncu_gemm_ff2.log 2>&1; tail -n 1 gpurun_out/ncu_gemm_ff2.log
cp generativemodels_b200/csrc/.obj/igemm.o gpurun_out/igemm_profiled.o
