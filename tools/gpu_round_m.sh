#!/bin/bash
mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests/test_kernels_gpu.py -m gpu -q -p no:cacheprovider -x > gpurun_out/km.log 2>&1; echo "kernel tests exit $?"; tail -12 gpurun_out/km.log
timeout -k 10 900 python -m pytest tests/test_parity_gpu.py tests/test_fullsize_gpu.py -m gpu -q -p no:cacheprovider -x > gpurun_out/pm.log 2>&1; echo "parity+fullsize tests exit $?"; tail -5 gpurun_out/pm.log
timeout -k 10 600 python tools/perf_c3.py > gpurun_out/perf_m.log 2>&1; head -45 gpurun_out/perf_m.log | tail -40
