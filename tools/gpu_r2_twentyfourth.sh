#!/bin/bash
# ncu --set full of the CTA-pair 3x3x3 convolution at HEAD (the epilogue changed in the second half of round 2: 155 -> ~190
# registers), full-resolution 256 -> 256 call of one C3 forward.
mkdir -p gpurun_out
timeout -k 10 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
  -k regex:"igemm_tc_kernel<\(int\)256" -s 1 -c 1 -f -o gpurun_out/prof_igemm_final3 \
  python tools/perf_c3.py --shape 160,224,160 --iters 0 --breakdown 0 > gpurun_out/ncu_igemm_final3.log 2>&1
tail -n 2 gpurun_out/ncu_igemm_final3.log
ncu -i gpurun_out/prof_igemm_final3.ncu-rep --page details > gpurun_out/prof_igemm_final3_details.txt 2>/dev/null
grep -E "Duration|SM Frequency|Registers Per|Executed Ipc Active|DRAM Throughput|L2 Cache Throughput" gpurun_out/prof_igemm_final3_details.txt | head
ncu -i gpurun_out/prof_igemm_final3.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv, sys
rows = list(csv.reader(sys.stdin))
h = rows[0]
for name in ('sm__inst_executed_pipe_tensor_op_hmma.avg.pct_of_peak_sustained_active','sm__pipe_tensor_op_hmma_cycles_active.avg.pct_of_peak_sustained_active','sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active','dram__bytes_read.sum','dram__bytes_write.sum','gpu__time_duration.sum'):
    for i, n in enumerate(h):
        if n == name: print(name, rows[2][i], rows[1][i])
for i, n in enumerate(h):
    if 'tensor' in n and 'pct' in n: print(n, rows[2][i])
" | head -20
rm -f gpurun_out/prof_igemm_final3.ncu-rep
