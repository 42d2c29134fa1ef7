#!/bin/bash
mkdir -p gpurun_out
python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; tail -c 400 gpurun_out/bench_default.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_default.json"))
print(round(d["value"]), "evals/s e2e", round(d["e2e"]["value"]), {k: round(x, 3) for k, x in d["stage_ms"].items()}, "roofline frac", round(d["roofline"]["frac"], 3))
for k, v in d["configs"].items():
    print(k, json.dumps(v)[:600])
PY
