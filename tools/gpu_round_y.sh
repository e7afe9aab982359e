#!/bin/bash
mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests/test_kernels_gpu.py tests/test_parity_gpu.py tests/test_reference_suite_gpu.py -m gpu -q -p no:cacheprovider -x -k "causal or embed or transformer or rows_linear or attention_decode" > gpurun_out/kv.log 2>&1; echo "transformer tests exit $?"; tail -12 gpurun_out/kv.log
timeout -k 10 600 python tools/decode_probe.py > gpurun_out/decode.log 2>&1; cat gpurun_out/decode.log | tail -8
