#!/bin/bash
# Register-cached single-launch GroupNorm, out-conv GEMM padded to 32 columns, gn_finalize with loads in flight; the
# "stop narrowing at one wave" rule for long reductions as an experiment (B200_NARROW_ONE_WAVE).
mkdir -p gpurun_out
timeout -k 10 600 python -m pytest tests/test_kernels_gpu.py -q -p no:cacheprovider -x 2>&1 | tail -n 2
for w in 0 1; do
  echo "== B200_NARROW_ONE_WAVE=$w"
  B200_NARROW_ONE_WAVE=$w timeout 600 python tools/gemm_probe.py --split-ab 2>&1 | tail -n 15 | cut -c1-60
  B200_NARROW_ONE_WAVE=$w timeout 900 python tools/splitk_ab.py _SPLIT_K 2>&1 | tail -n 4
done
timeout -k 10 1500 python -m pytest tests -m gpu -q -p no:cacheprovider -x > gpurun_out/suite_twentieth.log 2>&1; echo "suite exit $?"; tail -n 2 gpurun_out/suite_twentieth.log
