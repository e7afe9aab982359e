#!/bin/bash
# what the driver runs at round end, in one call: GPU tests, smoke, both bench arms
mkdir -p gpurun_out
timeout -k 10 1500 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/all_gpu_tests.log 2>&1; echo "gpu tests exit $?"; tail -3 gpurun_out/all_gpu_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
timeout -k 10 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.log 2>&1; tail -1 gpurun_out/bench_ref.log | cut -c1-200
timeout -k 10 1500 python bench.py > gpurun_out/bench.log 2>&1; tail -1 gpurun_out/bench.log
