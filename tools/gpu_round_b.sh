#!/bin/bash
mkdir -p gpurun_out
export CUDA_LAUNCH_BLOCKING=1
timeout -k 10 300 python -m pytest tests/test_kernels_gpu.py -m gpu -q -p no:cacheprovider -k "tensorcore or rescale" > gpurun_out/kb.log 2>&1; echo "attn kernels exit $?"; grep -E "passed|failed|FAILED|rel-L2" gpurun_out/kb.log | head -20
unset CUDA_LAUNCH_BLOCKING
timeout -k 10 600 python -m pytest tests/test_parity_gpu.py -m gpu -q -p no:cacheprovider > gpurun_out/pb.log 2>&1; echo "parity exit $?"; tail -4 gpurun_out/pb.log
timeout -k 10 900 python tools/perf_c3.py --shape 160,224,160 --iters 2 > gpurun_out/perf_c3.log 2>&1; tail -14 gpurun_out/perf_c3.log
