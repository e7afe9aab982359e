#!/bin/bash
mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests/test_kernels_gpu.py -m gpu -q -p no:cacheprovider -x > gpurun_out/kn.log 2>&1; echo "kernel tests exit $?"; tail -12 gpurun_out/kn.log
timeout -k 10 900 python -m pytest tests/test_parity_gpu.py tests/test_fullsize_gpu.py -m gpu -q -p no:cacheprovider -x > gpurun_out/pn.log 2>&1; echo "parity+fullsize tests exit $?"; tail -5 gpurun_out/pn.log
timeout -k 10 600 python tools/perf_c3.py --gnfuse 0 > gpurun_out/perf_n0.log 2>&1; head -16 gpurun_out/perf_n0.log | tail -11
timeout -k 10 600 python tools/perf_c3.py --gnfuse 1 > gpurun_out/perf_n1.log 2>&1; head -22 gpurun_out/perf_n1.log | tail -17
