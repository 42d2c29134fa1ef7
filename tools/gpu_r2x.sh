#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-configs > gpurun_out/bench_short.json 2> gpurun_out/bench_short.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_short.json"))
print(round(d["value"]), "evals/s e2e", round(d["e2e"]["value"]), "pageable", round(d["e2e"]["pageable_frame"]["value"]), {k: round(x, 3) for k, x in d["stage_ms"].items()}, d["gpu_launches"])
PY
