#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -p no:cacheprovider > gpurun_out/split2_tests.log 2>&1; echo "gpu tests exit $?"
tail -4 gpurun_out/split2_tests.log; grep -E "^(FAILED|ERROR)" gpurun_out/split2_tests.log | head
timeout 600 python tools/brain_ldm_probe.py > gpurun_out/brain_ldm_split2.log 2>&1; echo "probe exit $?"
tail -9 gpurun_out/brain_ldm_split2.log
timeout 900 python tools/run_configs.py > gpurun_out/configs_split2.log 2>&1; echo "configs exit $?"
tail -12 gpurun_out/configs_split2.log
timeout 300 python tools/abi_breakdown.py brain > gpurun_out/breakdown_brain2.log 2>&1; head -45 gpurun_out/breakdown_brain2.log
