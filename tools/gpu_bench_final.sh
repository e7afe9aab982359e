#!/bin/bash
mkdir -p gpurun_out
timeout -k 10 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -p no:cacheprovider -k "upsample" > gpurun_out/ku.log 2>&1; echo "upsample exit $?"; tail -2 gpurun_out/ku.log
timeout -k 10 600 python -m pytest tests/test_parity_gpu.py -m gpu -q -p no:cacheprovider > gpurun_out/pf.log 2>&1; echo "parity exit $?"; tail -2 gpurun_out/pf.log
timeout -k 10 1500 python bench.py > gpurun_out/bench.log 2>&1; tail -1 gpurun_out/bench.log
timeout -k 10 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.log 2>&1; tail -1 gpurun_out/bench_ref.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
