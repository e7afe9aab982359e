#!/bin/bash
mkdir -p gpurun_out
timeout -k 10 900 ncu --set full --clock-control none --import-source on -k regex:flash512 -c 1 -f -o gpurun_out/prof_flash2p \
  python tools/attn_probe.py replay > gpurun_out/ncu_flash2p.log 2>&1
tail -3 gpurun_out/ncu_flash2p.log; ls -la gpurun_out/prof_flash2p.ncu-rep
