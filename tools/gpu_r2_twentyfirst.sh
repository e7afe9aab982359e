#!/bin/bash
# 4-channel GroupNorm partial groups from the conv epilogue (128-channel tensors), GEMV for a handful of context tokens:
# kernel tests, suite, graph-replayed UNet steps.
mkdir -p gpurun_out
timeout -k 10 600 python -m pytest tests/test_kernels_gpu.py -q -p no:cacheprovider -x 2>&1 | tail -n 3
timeout -k 10 1500 python -m pytest tests -m gpu -q -p no:cacheprovider -x > gpurun_out/suite_twentyfirst.log 2>&1; echo "suite exit $?"; tail -n 3 gpurun_out/suite_twentyfirst.log
timeout 900 python tools/splitk_ab.py _SPLIT_K 2>&1 | tail -n 4
