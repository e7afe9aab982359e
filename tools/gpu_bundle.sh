#!/bin/bash
# bundle (rank 4) GPU check: its tests + the full-size probe
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_bundle.py -x -q -m gpu > gpurun_out/bundle_tests.log 2>&1; echo "bundle tests exit $?"
tail -5 gpurun_out/bundle_tests.log
timeout 900 python tools/brain_ldm_probe.py > gpurun_out/brain_ldm.log 2>&1; echo "probe exit $?"
tail -30 gpurun_out/brain_ldm.log
