#!/bin/bash
mkdir -p gpurun_out
timeout -k 10 600 python -m pytest tests/test_parity_gpu.py -m gpu -q -p no:cacheprovider -k "cuda_graph" > gpurun_out/pg.log 2>&1; echo "graph test exit $?"; tail -15 gpurun_out/pg.log
timeout -k 10 1200 python tools/run_configs.py > gpurun_out/configs.log 2>&1; tail -10 gpurun_out/configs.log
