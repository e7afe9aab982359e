#!/bin/bash
# A/B of (a) the cooperative one-launch split-K and (b) programmatic dependent launch with the late trigger, in one box:
# B200_PDL in {0, 2, 1} x _SPLIT_FUSED in {False, True}, graph-replayed brain-LDM / C2 / C5 UNet steps.
mkdir -p gpurun_out
timeout -k 10 300 python -m pytest tests/test_kernels_gpu.py -q -p no:cacheprovider -k "split_k" 2>&1 | tail -n 3
for pdl in 0 2 1; do
  echo "== B200_PDL=$pdl"
  B200_PDL=$pdl timeout 600 python tools/splitk_ab.py _SPLIT_FUSED 2>&1 | tail -n 3
done
B200_PDL=2 timeout -k 10 600 python -m pytest tests -m gpu -q -p no:cacheprovider -x -k "not fullsize" > gpurun_out/suite_pdl2.log 2>&1; echo "suite (B200_PDL=2) exit $?"; tail -n 2 gpurun_out/suite_pdl2.log
