#!/bin/bash
mkdir -p gpurun_out
timeout -k 10 600 python tools/perf_c3.py --shape 64,64,64 --iters 1 > gpurun_out/perf_small.log 2>&1; tail -16 gpurun_out/perf_small.log
timeout -k 10 900 python tools/perf_c3.py --shape 160,224,160 --iters 2 > gpurun_out/perf_c3.log 2>&1; tail -20 gpurun_out/perf_c3.log
timeout -k 10 1200 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.log 2>&1; tail -5 gpurun_out/bench.log
