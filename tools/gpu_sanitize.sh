#!/bin/bash
# compute-sanitizer over the small-shape GPU tests: memcheck (kernel + parity tests), racecheck and synccheck (kernel tests)
mkdir -p gpurun_out
K="not fullres and not flash_replay_matches and not perf"
timeout -k 10 900 compute-sanitizer --tool memcheck --error-exitcode 7 --print-limit 20 \
  python -m pytest tests/test_kernels_gpu.py tests/test_parity_gpu.py -m gpu -q -p no:cacheprovider -x -k "$K" > gpurun_out/sanitize_mem.log 2>&1
echo "memcheck exit $?"; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/sanitize_mem.log | tail -3
timeout -k 10 900 compute-sanitizer --tool racecheck --error-exitcode 7 --print-limit 20 \
  python -m pytest tests/test_kernels_gpu.py -m gpu -q -p no:cacheprovider -x -k "$K" > gpurun_out/sanitize_race.log 2>&1
echo "racecheck exit $?"; grep -E "RACECHECK SUMMARY|ERROR SUMMARY|hazard|passed|failed" gpurun_out/sanitize_race.log | tail -6
timeout -k 10 900 compute-sanitizer --tool synccheck --error-exitcode 7 --print-limit 20 \
  python -m pytest tests/test_kernels_gpu.py -m gpu -q -p no:cacheprovider -x -k "$K" > gpurun_out/sanitize_sync.log 2>&1
echo "synccheck exit $?"; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/sanitize_sync.log | tail -3
