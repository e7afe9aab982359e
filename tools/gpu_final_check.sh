#!/bin/bash
# what the driver runs at round end, in one call: GPU tests, smoke, both bench arms (+ the per-entry-point breakdown
# of the two latency-bound UNets)
mkdir -p gpurun_out
timeout -k 10 1500 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/all_gpu_tests.log 2>&1; rc=$?
echo "gpu tests exit $rc"; tail -3 gpurun_out/all_gpu_tests.log
if [ $rc -ne 0 ]; then
  grep -E "^(FAILED|ERROR)" gpurun_out/all_gpu_tests.log | head -20
  B200_SPLIT_K=0 timeout -k 10 1500 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/all_gpu_tests_nosplit.log 2>&1
  echo "gpu tests with B200_SPLIT_K=0 exit $?"; tail -3 gpurun_out/all_gpu_tests_nosplit.log
  grep -E "^(FAILED|ERROR)" gpurun_out/all_gpu_tests_nosplit.log | head -20
fi
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
timeout -k 10 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.log 2>&1; tail -1 gpurun_out/bench_ref.log | cut -c1-200
timeout -k 10 1500 python bench.py > gpurun_out/bench.log 2>&1; tail -1 gpurun_out/bench.log
for w in brain c2; do
  timeout 300 python tools/abi_breakdown.py $w > gpurun_out/breakdown_$w.log 2>&1; echo "breakdown $w exit $?"
  B200_SPLIT_K=0 timeout 300 python tools/abi_breakdown.py $w > gpurun_out/breakdown_${w}_nosplit.log 2>&1
  head -14 gpurun_out/breakdown_$w.log; head -1 gpurun_out/breakdown_${w}_nosplit.log
done
