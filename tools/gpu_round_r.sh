#!/bin/bash
mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests/test_kernels_gpu.py -m gpu -q -p no:cacheprovider -x -k "attention" > gpurun_out/ka.log 2>&1; echo "attention tests exit $?"; tail -3 gpurun_out/ka.log
timeout -k 10 300 python tools/attn_probe.py replay > gpurun_out/attn_r.log 2>&1; cat gpurun_out/attn_r.log
timeout -k 10 1200 python -m pytest tests/test_parity_gpu.py tests/test_kernels_gpu.py -m gpu -q -p no:cacheprovider -x > gpurun_out/pr.log 2>&1; echo "parity+kernel tests exit $?"; tail -6 gpurun_out/pr.log
