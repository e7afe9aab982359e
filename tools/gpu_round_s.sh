#!/bin/bash
mkdir -p gpurun_out
timeout -k 10 1200 python -m pytest tests/test_kernels_gpu.py tests/test_parity_gpu.py -m gpu -q -p no:cacheprovider -x > gpurun_out/ks.log 2>&1; echo "kernel+parity tests exit $?"; tail -4 gpurun_out/ks.log
timeout -k 10 600 python tools/perf_c3.py > gpurun_out/perf_s.log 2>&1; head -45 gpurun_out/perf_s.log | tail -40
