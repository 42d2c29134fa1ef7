#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "search_matches or ragged or point_size or overflow or retries or sequence or relocalize" > gpurun_out/pytest_r.log 2>&1; tail -2 gpurun_out/pytest_r.log
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-configs > gpurun_out/bench_short.json 2> gpurun_out/bench_short.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_short.json"))
print(round(d["value"]), "evals/s e2e", round(d["e2e"]["value"]), {k: round(x, 3) for k, x in d["stage_ms"].items()})
PY
python tools/exp_overflow_retry.py 2>&1 | tail -6
