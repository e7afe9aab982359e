#!/bin/bash
# One-launch (cooperative) split-K under the final planner (splits only with >= 3 ranges of >= 32 chunks): A/B on the
# graph-replayed UNet steps, then the whole suite with it on.
mkdir -p gpurun_out
timeout 600 python tools/splitk_ab.py _SPLIT_FUSED 2>&1 | tail -n 4
B200_SPLIT_FUSED=1 timeout -k 10 900 python -m pytest tests -m gpu -q -p no:cacheprovider -x > gpurun_out/suite_fused_final.log 2>&1; echo "suite (B200_SPLIT_FUSED=1) exit $?"; tail -n 2 gpurun_out/suite_fused_final.log
