#!/bin/bash
# compute-sanitizer (memcheck, synccheck) over the kernel tests that reach the code of round 2's second half: fused GEGLU,
# lean epilogue + bias prefetch + residual prefetch, batched V^T (a_broadcast), cooperative one-launch split-K, 4-channel
# GroupNorm partial groups, register-cached GroupNorm, 128-bit layernorm, tap reformulations.
mkdir -p gpurun_out
K="geglu or layernorm or split_k or groupnorm or linear or tap_reform or concat_epilogue or attention_tc"
timeout -k 10 900 compute-sanitizer --tool memcheck --error-exitcode 7 --print-limit 20 \
  python -m pytest tests/test_kernels_gpu.py -m gpu -q -p no:cacheprovider -x -k "$K" > gpurun_out/sanitize2_mem.log 2>&1
echo "memcheck exit $?"; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/sanitize2_mem.log | tail -3
timeout -k 10 900 compute-sanitizer --tool synccheck --error-exitcode 7 --print-limit 20 \
  python -m pytest tests/test_kernels_gpu.py -m gpu -q -p no:cacheprovider -x -k "$K" > gpurun_out/sanitize2_sync.log 2>&1
echo "synccheck exit $?"; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/sanitize2_sync.log | tail -3
