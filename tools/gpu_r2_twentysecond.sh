#!/bin/bash
# Epilogue experiments: next tile's bias vector prefetched (B200_BIAS_PREFETCH 1/0), 128-bit instead of 256-bit stores
# in the lean body (B200_NO_V256): kernel tests with the new default, GEMM table per setting, UNet steps per setting.
mkdir -p gpurun_out
timeout -k 10 600 python -m pytest tests/test_kernels_gpu.py -q -p no:cacheprovider -x 2>&1 | tail -n 2
for cfg in "B200_BIAS_PREFETCH=1 B200_NO_V256=0" "B200_BIAS_PREFETCH=0 B200_NO_V256=0" "B200_BIAS_PREFETCH=1 B200_NO_V256=1"; do
  echo "== $cfg"
  env $cfg timeout 600 python tools/gemm_probe.py 2>&1 | tail -n 15 | head -n 11
  env $cfg timeout 900 python tools/splitk_ab.py _SPLIT_K 2>&1 | tail -n 2
done
