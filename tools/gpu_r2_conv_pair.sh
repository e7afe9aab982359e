#!/bin/bash
mkdir -p gpurun_out
for pair in 0 1; do
  echo "== B200_IGEMM_PAIR=$pair"
  timeout -k 10 300 env B200_IGEMM_PAIR=$pair python tools/conv_probe.py 2>&1 | tail -n 3
done
echo "== full-size conv cross-check with the pair kernel"
timeout -k 10 600 env B200_IGEMM_PAIR=1 python -m pytest tests/test_fullsize_gpu.py tests/test_kernels_gpu.py -q -p no:cacheprovider -k "conv" 2>&1 | tail -n 4
