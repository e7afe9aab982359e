#!/bin/bash
mkdir -p gpurun_out
echo "== suite"
timeout -k 10 1500 python -m pytest tests -m gpu -q -p no:cacheprovider --maxfail=40 > gpurun_out/suite8.log 2>&1
echo "suite exit $?"; tail -n 3 gpurun_out/suite8.log; grep -E "^(FAILED|ERROR)" gpurun_out/suite8.log | head -40
echo "== other configs (128-column pair tiles)"
timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_v7.json 2>/dev/null
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_v7.json"))
print("C3 ms/step", round(d["ms_per_step"], 1), {k: (round(v["ms_per_call"], 2), v.get("ms_per_call_eager")) for k, v in d["other_configs"].items()})
PY
echo "== compute-sanitizer on the pair kernels"
K="pair or 160x128 or 151x129 or 104x128 or 12x40x40 or 300x260 or 100x128 or flash_rescale or attention_tensorcore"
timeout -k 10 900 compute-sanitizer --tool memcheck --error-exitcode 7 --print-limit 20 \
  python -m pytest tests/test_kernels_gpu.py -m gpu -q -p no:cacheprovider -x -k "$K" > gpurun_out/sanitize_mem_r2.log 2>&1
echo "memcheck exit $?"; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/sanitize_mem_r2.log | tail -3
timeout -k 10 900 compute-sanitizer --tool racecheck --error-exitcode 7 --print-limit 20 \
  python -m pytest tests/test_kernels_gpu.py -m gpu -q -p no:cacheprovider -x -k "$K" > gpurun_out/sanitize_race_r2.log 2>&1
echo "racecheck exit $?"; grep -E "RACECHECK SUMMARY|ERROR SUMMARY|hazard|passed|failed" gpurun_out/sanitize_race_r2.log | tail -6
timeout -k 10 900 compute-sanitizer --tool synccheck --error-exitcode 7 --print-limit 20 \
  python -m pytest tests/test_kernels_gpu.py -m gpu -q -p no:cacheprovider -x -k "$K" > gpurun_out/sanitize_sync_r2.log 2>&1
echo "synccheck exit $?"; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/sanitize_sync_r2.log | tail -3
