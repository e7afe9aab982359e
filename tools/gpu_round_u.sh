#!/bin/bash
mkdir -p gpurun_out
timeout -k 5 240 python -m pytest tests/test_kernels_gpu.py -m gpu -q -p no:cacheprovider -x -k "attention" > gpurun_out/ka.log 2>&1; echo "attention tests exit $?"; tail -5 gpurun_out/ka.log
timeout -k 5 120 python tools/attn_probe.py replay > gpurun_out/attn_u.log 2>&1; echo "probe exit $?"; cat gpurun_out/attn_u.log | tail -4
B200_FLASH_PAIR=0 timeout -k 5 120 python tools/attn_probe.py replay > gpurun_out/attn_u0.log 2>&1; echo "probe(no pair) exit $?"; cat gpurun_out/attn_u0.log | tail -4
