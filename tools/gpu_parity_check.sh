#!/bin/bash
mkdir -p gpurun_out
run() {
  timeout -k 10 $3 python -m pytest $1 -m gpu -q -p no:cacheprovider -x > gpurun_out/$2.log 2>&1
  echo "$2 exit $?"; tail -25 gpurun_out/$2.log
}
run tests/test_parity_gpu.py parity 900
run tests/test_kernels_gpu.py kernels 600
timeout -k 10 600 python tools/perf_c3.py --shape 64,64,64 > gpurun_out/perf_small.log 2>&1; tail -20 gpurun_out/perf_small.log
timeout -k 10 900 python tools/perf_c3.py --shape 160,224,160 > gpurun_out/perf_c3.log 2>&1; tail -25 gpurun_out/perf_c3.log
