#!/bin/bash
# GPU run: op-level parity of every kernel, split into independent processes so one sticky CUDA error or hang
# cannot hide the other groups.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
export CUDA_LAUNCH_BLOCKING=${CUDA_LAUNCH_BLOCKING:-1}
run() {  # name, -k expression
  timeout -k 10 300 python -m pytest tests/test_kernels_gpu.py -m gpu -q -s -p no:cacheprovider -k "$2" > gpurun_out/k_$1.log 2>&1
  echo "$1 exit $?" | tee -a gpurun_out/k_$1.log
  grep -E "passed|failed|perf-probe" gpurun_out/k_$1.log | tail -3
  grep -E "^(FAILED|ERROR)" gpurun_out/k_$1.log | head -12
}
run misc "layout or groupnorm or layernorm or resample or attention_small or timestep"
run check "check"
run tc_s1 "tcgen05 and s1 and not concat"
run tc_s2 "tcgen05 and s2"
run tc_epi "tcgen05 and concat"
run tr "asym or transpose"
run lin "linear"
run attn "tensorcore"
run perf "perf"
