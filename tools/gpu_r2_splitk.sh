#!/bin/bash
mkdir -p gpurun_out
timeout -k 10 300 python -m pytest tests/test_kernels_gpu.py -q -p no:cacheprovider -k "split_k" 2>&1 | tail -n 5
for i in 1 2; do timeout 600 python tools/splitk_ab.py _SPLIT_FUSED 2>&1 | tail -n 2; done
timeout -k 10 1500 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/suite_fused.log 2>&1; echo "suite exit $?"; tail -n 2 gpurun_out/suite_fused.log; grep -E "^(FAILED|ERROR)" gpurun_out/suite_fused.log | head
