#!/bin/bash
# round 2, run D: whole GPU suite (new tests included) + quick bench
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -25 gpurun_out/pytest_gpu.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_d.json 2> gpurun_out/bench_d.err
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/bench_d.json"))
    print(round(d["value"]), "evals/s e2e", round(d["e2e"]["value"]), {k: round(x, 3) for k, x in d["stage_ms"].items()}, "launches", d["gpu_launches"])
except Exception as e:
    print("FAILED", e, open("gpurun_out/bench_d.err").read()[-600:])
PY
