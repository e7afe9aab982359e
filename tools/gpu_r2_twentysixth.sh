#!/bin/bash
# Independent branches (V^T projection, 1x1 skip convolution) on a side stream inside CUDA-graph captures (B200_FORK 1/0):
# the whole suite (graph-vs-eager tests included) with the default, then graph-replayed UNet steps per setting.
mkdir -p gpurun_out
timeout -k 10 1500 python -m pytest tests -m gpu -q -p no:cacheprovider -x > gpurun_out/suite_twentysixth.log 2>&1; echo "suite exit $?"; tail -n 3 gpurun_out/suite_twentysixth.log
for v in 1 0; do
  echo "== B200_FORK=$v"
  B200_FORK=$v timeout 900 python tools/splitk_ab.py _SPLIT_K 2>&1 | tail -n 4
done
