#!/bin/bash
mkdir -p gpurun_out
B200_FLASH_PAIR=0 B200_DEV_LIB=generativemodels_b200/lib/dev/libb200gen_timing.so timeout -k 10 300 python tools/attn_timing.py > gpurun_out/attn_timing.log 2>&1; cat gpurun_out/attn_timing.log
B200_FLASH_PAIR=1 B200_DEV_LIB=generativemodels_b200/lib/dev/libb200gen_timing.so timeout -k 10 300 python tools/attn_timing.py > gpurun_out/attn_timing1.log 2>&1; cat gpurun_out/attn_timing1.log
