#!/bin/bash
# histogram contention cases (not the BASELINE workload): frame statistics x cloud extent x skip mode
mkdir -p gpurun_out
for cfg in ${CASES:-textured_40_0 textured_40_1 textured_40_2 sky_40_1 sky_12_1 constant_40_1}; do
set -- ${cfg//_/ }
NMI_HIST_SKIP=$3 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --frame $1 --extent $2 > gpurun_out/bench_stress.json 2> gpurun_out/bench_stress.err
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_stress.json"))
    print("$1 extent $2 skip $3:", round(d["value"]), "evals/s", {k: round(x,3) for k,x in d["stage_ms"].items() if k in ("render","hist_score","total")}, d["config"]["winner"])
except Exception as e:
    print("FAILED", e, open("gpurun_out/bench_stress.err").read()[-400:])
PY
done
