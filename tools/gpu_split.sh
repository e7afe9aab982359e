#!/bin/bash
# split-K check: whole GPU suite (the split path is taken by most small test shapes), then the latency-bound configs
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/split_tests.log 2>&1; echo "gpu tests exit $?"
tail -15 gpurun_out/split_tests.log
timeout 600 python tools/brain_ldm_probe.py > gpurun_out/brain_ldm_split.log 2>&1; echo "probe exit $?"
tail -12 gpurun_out/brain_ldm_split.log
timeout 900 python tools/run_configs.py > gpurun_out/configs_split.log 2>&1; echo "configs exit $?"
tail -20 gpurun_out/configs_split.log
