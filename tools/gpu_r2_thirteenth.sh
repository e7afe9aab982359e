#!/bin/bash
# Short-reduction GEMMs: timing table + ncu --set full captures (source-level) of two epilogue-exposed shapes.
mkdir -p gpurun_out
timeout 600 python tools/gemm_probe.py 2>&1 | tail -n 20
timeout -k 10 600 ncu --profile-from-start off --set full --clock-control none --import-source on -f -o gpurun_out/prof_gemm_ff \
  python tools/gemm_probe.py --profile 32768,2048,256 > gpurun_out/ncu_gemm_ff.log 2>&1; tail -n 2 gpurun_out/ncu_gemm_ff.log
timeout -k 10 600 ncu --profile-from-start off --set full --clock-control none --import-source on -f -o gpurun_out/prof_gemm_small \
  python tools/gemm_probe.py --profile 8192,256,256,1 > gpurun_out/ncu_gemm_small.log 2>&1; tail -n 2 gpurun_out/ncu_gemm_small.log
ls -la gpurun_out/*.ncu-rep
