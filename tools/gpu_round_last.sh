#!/bin/bash
mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests/test_parity_gpu.py tests/test_kernels_gpu.py -m gpu -q -p no:cacheprovider -x > gpurun_out/last.log 2>&1; echo "parity+kernel tests exit $?"; tail -3 gpurun_out/last.log
timeout -k 10 200 python tools/sampling_probe.py 2>&1 | tail -3
timeout -k 10 600 python tools/run_configs.py 2>&1 | grep -E "C2 LDM DDIM-50 N=1|transformer sampler 1024" 
