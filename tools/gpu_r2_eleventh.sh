#!/bin/bash
# A/B of the cooperative one-launch split-K with eight ranges' loads in flight (second form), same box, same process.
mkdir -p gpurun_out
timeout -k 10 300 python -m pytest tests/test_kernels_gpu.py -q -p no:cacheprovider -k "split_k" 2>&1 | tail -n 3
for i in 1 2; do timeout 600 python tools/splitk_ab.py _SPLIT_FUSED 2>&1 | tail -n 3; done
