#!/bin/bash
# quick A/B: bench only (no tests); ENVS="A=1;B=2" style variants separated by ';'
mkdir -p gpurun_out
IFS=';' read -ra VARS <<< "${ENVS:-X=0}"
for v in "${VARS[@]}"; do
env $v timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_ab.json 2> gpurun_out/bench_ab.err
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_ab.json"))
    print("$v", round(d["value"]), "evals/s e2e", round(d["e2e"]["value"]), {k: round(x,3) for k,x in d["stage_ms"].items()}, "launches", d["gpu_launches"])
except Exception as e:
    print("FAILED", e, open("gpurun_out/bench_ab.err").read()[-400:])
PY
done
