#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
python tools/exp_small_grid_slice.py 2>&1 | tail -4
NMI_STREAM_PRIO=0 python tools/exp_small_grid_slice.py 2>&1 | tail -1
NMI_CULL_PASSES=3 python tools/exp_small_grid_slice.py 2>&1 | tail -1
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-configs > gpurun_out/bench_short.json 2> gpurun_out/bench_short.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_short.json"))
print(round(d["value"]), "evals/s e2e", round(d["e2e"]["value"]), {k: round(x, 3) for k, x in d["stage_ms"].items()})
PY
