#!/bin/bash
mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests/test_kernels_gpu.py -m gpu -q -p no:cacheprovider -x -k "attention" > gpurun_out/ka.log 2>&1; echo "attention tests exit $?"; tail -4 gpurun_out/ka.log
timeout -k 10 600 python tools/attn_probe.py > gpurun_out/attn_l.log 2>&1; cat gpurun_out/attn_l.log
