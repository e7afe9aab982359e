#!/bin/bash
mkdir -p gpurun_out
timeout -k 10 1200 python -m pytest tests/test_kernels_gpu.py tests/test_parity_gpu.py -m gpu -q -p no:cacheprovider -x > gpurun_out/kj.log 2>&1; echo "kernel+parity tests exit $?"; tail -8 gpurun_out/kj.log
timeout -k 10 600 python tools/perf_c3.py > gpurun_out/perf_j.log 2>&1; tail -42 gpurun_out/perf_j.log
