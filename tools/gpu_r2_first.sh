#!/bin/bash
# Round-2 first GPU call: the suite on the fp16 default, the staged (round-1, never run on hardware) features, the bf16
# flavour, the headline bench and the other configurations' wall clocks.
mkdir -p gpurun_out
run() { name=$1; shift; timeout -k 10 1500 env "$@" python -m pytest tests -m gpu -q -p no:cacheprovider --maxfail=40 > gpurun_out/$name.log 2>&1
        echo "$name exit $?"; tail -3 gpurun_out/$name.log; grep -E "^(FAILED|ERROR)" gpurun_out/$name.log | head -40; }
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv,noheader
run suite_head B200_NOOP=1
grep -E "^\[C[2345]\]|\[C2|perf-probe" gpurun_out/suite_head.log | head -20
timeout 300 env B200_STAGED=1 python -m pytest tests/test_kernels_gpu.py -q -k groupnorm_fused_small -p no:cacheprovider \
  > gpurun_out/staged_gn.log 2>&1; echo "staged GroupNorm tests exit $?"; tail -3 gpurun_out/staged_gn.log
run suite_gn_small B200_GN_SMALL=1 B200_STAGED=1
run suite_auto_graph B200_AUTO_GRAPH=1
timeout 900 env B200_ACT_DTYPE=bf16 python -m pytest tests -m gpu -q -p no:cacheprovider --maxfail=40 -k "not c2_reference and not fullsize" > gpurun_out/suite_bf16.log 2>&1
echo "suite_bf16 exit $?"; tail -3 gpurun_out/suite_bf16.log; grep -E "^(FAILED|ERROR)" gpurun_out/suite_bf16.log | head -20
timeout 300 python tools/splitk_ab.py _GN_SMALL > gpurun_out/gn_small_ab.log 2>&1; tail -3 gpurun_out/gn_small_ab.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_fp16.json 2> gpurun_out/bench_fp16.err; echo "bench exit $?"; cat gpurun_out/bench_fp16.json | cut -c1-600
timeout 900 python tools/run_configs.py > gpurun_out/other_configs.log 2>&1; echo "run_configs exit $?"; head -12 gpurun_out/other_configs.log
