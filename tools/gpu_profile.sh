#!/bin/bash
mkdir -p gpurun_out
timeout -k 10 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
  -k regex:"igemm_tc_kernel<256" -s 3 -c 1 -f -o gpurun_out/prof_igemm \
  python tools/perf_c3.py --shape 160,224,160 --iters 0 --breakdown 0 > gpurun_out/ncu_igemm.log 2>&1
tail -2 gpurun_out/ncu_igemm.log
timeout -k 10 900 ncu --set full --clock-control none --import-source on -k regex:flash_attn -c 1 -f -o gpurun_out/prof_flash \
  python tools/perf_c3.py --shape 160,224,160 --iters 0 --breakdown 0 > gpurun_out/ncu_flash.log 2>&1
tail -2 gpurun_out/ncu_flash.log
ls -la gpurun_out/*.ncu-rep
timeout -k 10 1200 python bench.py > gpurun_out/bench.log 2>&1; tail -2 gpurun_out/bench.log
