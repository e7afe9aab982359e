#!/bin/bash
mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests/test_reference_suite_gpu.py -m gpu -q -p no:cacheprovider -x > gpurun_out/rw.log 2>&1; echo "reference-suite tests exit $?"; tail -25 gpurun_out/rw.log
