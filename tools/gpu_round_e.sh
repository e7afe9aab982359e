#!/bin/bash
mkdir -p gpurun_out
timeout -k 10 400 python -m pytest tests/test_kernels_gpu.py -m gpu -q -p no:cacheprovider -x -k "tensorcore or rescale" > gpurun_out/ke.log 2>&1; echo "attn exit $?"; tail -3 gpurun_out/ke.log
timeout -k 10 900 python tools/perf_c3.py --shape 160,224,160 --iters 2 > gpurun_out/perf_c3.log 2>&1; tail -13 gpurun_out/perf_c3.log
timeout -k 10 900 python -m pytest tests/test_fullsize_gpu.py -m gpu -q -p no:cacheprovider > gpurun_out/fs.log 2>&1; echo "fullsize exit $?"; tail -8 gpurun_out/fs.log
timeout -k 10 600 python -m pytest tests/test_parity_gpu.py -m gpu -q -p no:cacheprovider -x > gpurun_out/pe.log 2>&1; echo "parity exit $?"; tail -3 gpurun_out/pe.log
