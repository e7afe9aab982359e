#!/bin/bash
# After (a) the batched V^T projection (a_broadcast), (b) batched bias loads + residual prefetch in the GEMM epilogue:
# kernel tests, the short-K GEMM table, graph-replayed UNet steps (brain / C2 / C2 batch 32 / C5), then the whole suite.
mkdir -p gpurun_out
timeout -k 10 600 python -m pytest tests/test_kernels_gpu.py -q -p no:cacheprovider -x 2>&1 | tail -n 4
timeout 600 python tools/gemm_probe.py 2>&1 | tail -n 16
timeout 900 python tools/splitk_ab.py _SPLIT_FUSED 2>&1 | tail -n 4
timeout -k 10 1500 python -m pytest tests -m gpu -q -p no:cacheprovider -x > gpurun_out/suite_fourteenth.log 2>&1; echo "suite exit $?"; tail -n 3 gpurun_out/suite_fourteenth.log
