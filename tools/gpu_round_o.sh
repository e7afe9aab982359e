#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/attn_ab.log
for n in 1 2 3 4; do
  echo "== ablation $n (1: no exp, 2: 1/32 of QK MMAs, 3: 1/4 of PV MMAs, 4: no P slab store)" >> gpurun_out/attn_ab.log
  B200_DEV_LIB=generativemodels_b200/lib/ablate/libb200gen_ab$n.so timeout -k 10 300 python tools/attn_probe.py >> gpurun_out/attn_ab.log 2>&1
done
echo "== baseline" >> gpurun_out/attn_ab.log
timeout -k 10 300 python tools/attn_probe.py >> gpurun_out/attn_ab.log 2>&1
grep -v "^$" gpurun_out/attn_ab.log | grep -v "\] T=S=89600 d=512: .* ms.*\n" | awk '/==/ {print} /flash/ {c[$0]=1; print}' | grep -v "^flash.*: 2[5-9]\.[0-9]* ms, 6[0-9][0-9] TFLOP.*XX"
