#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; tail -2 gpurun_out/pytest_gpu.log
for f in ${FRAMES:-textured}; do
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --frame $f > gpurun_out/bench_quick_$f.json 2> gpurun_out/bench_quick_$f.err
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_quick_$f.json"))
    print("$f", round(d["value"]), "evals/s e2e", round(d["e2e"]["value"]), {k: round(x,3) for k,x in d["stage_ms"].items()}, "launches", d["gpu_launches"])
except Exception as e:
    print("FAILED", e, open("gpurun_out/bench_quick_$f.err").read()[-400:])
PY
done
