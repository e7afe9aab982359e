#!/bin/bash
mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests/test_kernels_gpu.py -m gpu -q -p no:cacheprovider -x -k "attention" > gpurun_out/ka.log 2>&1; echo "attention tests exit $?"; tail -12 gpurun_out/ka.log
timeout -k 10 600 python tools/perf_c3.py > gpurun_out/perf_k.log 2>&1; head -22 gpurun_out/perf_k.log | tail -16
