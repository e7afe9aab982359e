#!/bin/bash
mkdir -p gpurun_out
B200_DEV_LIB=generativemodels_b200/lib/dev/libb200gen_ab6.so timeout -k 10 200 python tools/attn_probe.py replay 2>&1 | tail -3
timeout -k 10 200 python tools/attn_probe.py replay 2>&1 | tail -3
