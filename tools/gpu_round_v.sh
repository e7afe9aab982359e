#!/bin/bash
mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests/test_kernels_gpu.py tests/test_parity_gpu.py -m gpu -q -p no:cacheprovider -x -k "causal or embed or transformer or attention_small" > gpurun_out/kv.log 2>&1; echo "transformer tests exit $?"; tail -25 gpurun_out/kv.log
