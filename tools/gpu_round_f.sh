#!/bin/bash
mkdir -p gpurun_out
timeout -k 10 600 python tools/attn_probe.py > gpurun_out/attn_probe.log 2>&1; tail -3 gpurun_out/attn_probe.log
timeout -k 10 900 ncu --set full --clock-control none --import-source on -k regex:flash_attn -s 1 -c 1 -f -o gpurun_out/prof_flash_v3 \
  python tools/attn_probe.py > gpurun_out/ncu_flash3.log 2>&1; tail -2 gpurun_out/ncu_flash3.log
timeout -k 10 900 python tools/perf_c3.py --shape 160,224,160 --iters 2 --breakdown 1 > gpurun_out/perf_c3.log 2>&1; tail -13 gpurun_out/perf_c3.log
timeout -k 10 1200 python tools/run_configs.py > gpurun_out/configs.log 2>&1; tail -6 gpurun_out/configs.log
