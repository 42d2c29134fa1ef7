#!/bin/bash
# round 2, run N: small-grid (3^6 per level, 3 levels) per-frame cost with the per-level host trace, N = 1
mkdir -p gpurun_out
python bench.py --steps 5 --warmup 3 --frames 20 --no-cpu-baseline 2>gpurun_out/bench_n.err | tail -1 > gpurun_out/bench_n.json
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_n.json").read())
print(d["value"], d["ms_per_step"])
print(json.dumps(d["configs"].get("C5_small_grid"), indent=1)[:3000])
PY
