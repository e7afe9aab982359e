#!/bin/bash
# Round evidence: tests, bench (both arms), smoke, ncu launch list of the bench command, full captures of the two top kernels.
mkdir -p gpurun_out
timeout -k 10 1500 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/all_gpu_tests.log 2>&1; echo "gpu tests exit $?"; tail -3 gpurun_out/all_gpu_tests.log
timeout -k 10 1500 python bench.py > gpurun_out/bench.log 2>&1; tail -1 gpurun_out/bench.log
timeout -k 10 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.log 2>&1; tail -1 gpurun_out/bench_ref.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
timeout -k 10 1500 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_bench.csv \
  python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
tail -2 gpurun_out/ncu_bench.log
timeout -k 10 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
  -k regex:"igemm_tc_kernel<\(int\)256" -s 1 -c 1 -f -o gpurun_out/prof_igemm \
  python tools/perf_c3.py --shape 160,224,160 --iters 0 --breakdown 0 > gpurun_out/ncu_igemm.log 2>&1
tail -2 gpurun_out/ncu_igemm.log
timeout -k 10 900 ncu --set full --clock-control none --import-source on -k regex:flash_attn -c 1 -f -o gpurun_out/prof_flash \
  python tools/perf_c3.py --shape 160,224,160 --iters 0 --breakdown 0 > gpurun_out/ncu_flash.log 2>&1
tail -2 gpurun_out/ncu_flash.log
ls -la gpurun_out/*.ncu-rep
