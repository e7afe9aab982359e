#!/bin/bash
mkdir -p gpurun_out
timeout -k 10 400 python -m pytest tests/test_kernels_gpu.py -m gpu -q -p no:cacheprovider -k "tensorcore or rescale or groupnorm" > gpurun_out/kc.log 2>&1; echo "kernels exit $?"; tail -3 gpurun_out/kc.log
timeout -k 10 600 python -m pytest tests/test_parity_gpu.py -m gpu -q -p no:cacheprovider > gpurun_out/pc.log 2>&1; echo "parity exit $?"; tail -3 gpurun_out/pc.log
timeout -k 10 900 python tools/perf_c3.py --shape 160,224,160 --iters 2 > gpurun_out/perf_c3.log 2>&1; tail -13 gpurun_out/perf_c3.log
timeout -k 10 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
  -k regex:"igemm_tc_kernel<\(int\)256" -s 3 -c 1 -f -o gpurun_out/prof_igemm \
  python tools/perf_c3.py --shape 160,224,160 --iters 0 --breakdown 0 > gpurun_out/ncu_igemm.log 2>&1
tail -2 gpurun_out/ncu_igemm.log; ls -la gpurun_out/*.ncu-rep
