#!/bin/bash
mkdir -p gpurun_out
python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-configs > gpurun_out/bench_short.json 2> gpurun_out/bench_short.err; tail -c 300 gpurun_out/bench_short.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_short.json"))
print(round(d["value"]), "evals/s e2e", round(d["e2e"]["value"]), "pageable", round(d["e2e"]["pageable_frame"]["value"]), {k: round(x, 3) for k, x in d["stage_ms"].items()}, d["gpu_launches"], d["result"])
PY
