#!/bin/bash
# run-to-run spread of the histogram stage for a few kernel variants (same box, fresh process each)
mkdir -p gpurun_out
for v in ${VARIANTS:-0 8 0 8 0}; do
timeout 300 python bench.py --steps ${STEPS:-20} --warmup 3 --no-cpu-baseline --variant $v > gpurun_out/bench_var.json 2> gpurun_out/bench_var.err
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_var.json"))
    print("variant $v", round(d["value"]), "evals/s", {k: round(x,3) for k,x in d["stage_ms"].items() if k in ("hist_score","total")}, d["clocks"])
except Exception as e:
    print("FAILED", e, open("gpurun_out/bench_var.err").read()[-600:])
PY
done
