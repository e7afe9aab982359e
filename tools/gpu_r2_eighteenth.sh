#!/bin/bash
# Fused GEGLU feed-forward (linear1 + gating in one GEMM epilogue): kernel tests, the whole suite, graph-replayed UNet steps.
mkdir -p gpurun_out
timeout -k 10 600 python -m pytest tests/test_kernels_gpu.py -q -p no:cacheprovider -x 2>&1 | tail -n 3
timeout -k 10 1500 python -m pytest tests -m gpu -q -p no:cacheprovider -x > gpurun_out/suite_eighteenth.log 2>&1; echo "suite exit $?"; tail -n 3 gpurun_out/suite_eighteenth.log
timeout 900 python tools/splitk_ab.py _SPLIT_K 2>&1 | tail -n 4
