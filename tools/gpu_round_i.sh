#!/bin/bash
mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests/test_parity_gpu.py tests/test_reference_suite_gpu.py -m gpu -q -p no:cacheprovider -k "likelihood or kl or inferer" > gpurun_out/lik.log 2>&1; echo "likelihood tests exit $?"; tail -15 gpurun_out/lik.log
timeout -k 10 600 python tools/perf_c3.py > gpurun_out/perf_i.log 2>&1; tail -45 gpurun_out/perf_i.log
