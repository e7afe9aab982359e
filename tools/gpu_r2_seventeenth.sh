#!/bin/bash
# Coalesced (shared-memory transposed) lean epilogue + split range default 32 + 128-bit layernorm: kernel tests, GEMM
# table, graph-replayed UNet steps, the whole suite.
mkdir -p gpurun_out
timeout -k 10 600 python -m pytest tests/test_kernels_gpu.py -q -p no:cacheprovider -x 2>&1 | tail -n 3
timeout 600 python tools/gemm_probe.py 2>&1 | tail -n 16
timeout 900 python tools/splitk_ab.py _SPLIT_K 2>&1 | tail -n 4
timeout -k 10 1500 python -m pytest tests -m gpu -q -p no:cacheprovider -x > gpurun_out/suite_seventeenth.log 2>&1; echo "suite exit $?"; tail -n 3 gpurun_out/suite_seventeenth.log
