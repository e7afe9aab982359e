#!/bin/bash
# Lean GEMM epilogue (no run-time switches) + planner experiments: kernel tests, the short-K GEMM table, split-K per
# shape (one pass vs split), the narrowing rule for K = 256, graph-replayed UNet steps with the candidate settings.
mkdir -p gpurun_out
timeout -k 10 600 python -m pytest tests/test_kernels_gpu.py -q -p no:cacheprovider -x 2>&1 | tail -n 2
timeout 600 python tools/gemm_probe.py 2>&1 | tail -n 16
echo "== narrow from 4 chunks"; B200_NARROW_MIN_CHUNKS=4 timeout 600 python tools/gemm_probe.py 2>&1 | tail -n 16 | grep -E "M=8192|M=1024|M=256 "
echo "== split A/B"; timeout 600 python tools/gemm_probe.py --split-ab 2>&1 | tail -n 16
echo "== UNet steps, defaults"; timeout 900 python tools/splitk_ab.py _SPLIT_K 2>&1 | tail -n 4
echo "== UNet steps, B200_SPLIT_MIN=3"; B200_SPLIT_MIN=3 timeout 900 python tools/splitk_ab.py _SPLIT_K 2>&1 | tail -n 4
echo "== UNet steps, B200_SPLIT_MIN=4 B200_NARROW_MIN_CHUNKS=4"; B200_SPLIT_MIN=4 B200_NARROW_MIN_CHUNKS=4 timeout 900 python tools/splitk_ab.py _SPLIT_K 2>&1 | tail -n 4
