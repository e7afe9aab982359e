#!/bin/bash
# Round-2 second GPU call: suites (fp16 default, staged features on, bf16 flavour), fp16-vs-bf16 headline A/B on one box,
# the new bench line.
mkdir -p gpurun_out
run() { name=$1; shift; timeout -k 10 1500 env "$@" python -m pytest tests -m gpu -q -p no:cacheprovider --maxfail=40 > gpurun_out/$name.log 2>&1
        echo "$name exit $?"; tail -n 3 gpurun_out/$name.log; grep -E "^(FAILED|ERROR)" gpurun_out/$name.log | head -40; }
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv,noheader
run suite_head B200_NOOP=1
grep -E "^\[C[2345]\]|perf-probe" gpurun_out/suite_head.log | head -20
run suite_gn_small B200_GN_SMALL=1 B200_STAGED=1
run suite_auto_graph B200_AUTO_GRAPH=1
timeout 900 env B200_ACT_DTYPE=bf16 python -m pytest tests -m gpu -q -p no:cacheprovider --maxfail=40 -k "not c2_reference and not fullsize" > gpurun_out/suite_bf16.log 2>&1
echo "suite_bf16 exit $?"; tail -n 3 gpurun_out/suite_bf16.log; grep -E "^(FAILED|ERROR)" gpurun_out/suite_bf16.log | head -20
for rep in 1 2; do
  for dt in fp16 bf16; do
    timeout 600 env B200_ACT_DTYPE=$dt python bench.py --steps 6 --warmup 3 --no-other-configs --no-cpu-baseline > gpurun_out/ab_${dt}_$rep.json 2> gpurun_out/ab_${dt}_$rep.err
    python - <<PY
import json
d = json.load(open("gpurun_out/ab_${dt}_$rep.json"))
print("$dt rep $rep: ms/step", round(d["ms_per_step"], 1), "conv TF/s", round(d["roofline"]["achieved"], 1), "clock", d["clocks"]["sm_mhz"], [ (s["kernel"][:12], round(s["achieved"],1), round(s["frac"],3)) for s in d["roofline"]["secondary"]])
PY
  done
done
timeout 1200 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_v2.json 2> gpurun_out/bench_v2.err; echo "bench exit $?"; tail -n 5 gpurun_out/bench_v2.err; cut -c1-3000 gpurun_out/bench_v2.json
