#!/bin/bash
# First GPU call of the next round: re-verify HEAD, then the pieces that were committed without a GPU run
# (batched time-embedding projections are already on; single-launch GroupNorm and inferer auto-graph are off).
mkdir -p gpurun_out
run() { name=$1; shift; timeout -k 10 1200 env "$@" python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/$name.log 2>&1
        echo "$name exit $?"; tail -2 gpurun_out/$name.log; grep -E "^(FAILED|ERROR)" gpurun_out/$name.log | head -10; }
run suite_head B200_NOOP=1
timeout 300 env B200_STAGED=1 python -m pytest tests/test_kernels_gpu.py -q -k groupnorm_fused_small -p no:cacheprovider \
  > gpurun_out/staged_gn.log 2>&1; echo "staged GroupNorm tests exit $?"; tail -3 gpurun_out/staged_gn.log
run suite_gn_small B200_GN_SMALL=1
run suite_auto_graph B200_AUTO_GRAPH=1
timeout 300 python tools/splitk_ab.py _GN_SMALL > gpurun_out/gn_small_ab.log 2>&1; tail -3 gpurun_out/gn_small_ab.log
