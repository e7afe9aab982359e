#!/bin/bash
mkdir -p gpurun_out
timeout -k 10 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -p no:cacheprovider -x > gpurun_out/kd.log 2>&1; echo "kernels exit $?"; tail -4 gpurun_out/kd.log
timeout -k 10 600 python -m pytest tests/test_parity_gpu.py -m gpu -q -p no:cacheprovider -x > gpurun_out/pd.log 2>&1; echo "parity exit $?"; tail -3 gpurun_out/pd.log
timeout -k 10 900 python tools/perf_c3.py --shape 160,224,160 --iters 2 > gpurun_out/perf_c3.log 2>&1; tail -13 gpurun_out/perf_c3.log
