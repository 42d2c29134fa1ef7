#!/bin/bash
# round 2, run L: LDG (no shared-memory staging) variants 11-13 vs TMA variant 9 (one CTA per pair) and the default
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "flags_variants or overflow" > gpurun_out/pytest_l.log 2>&1; tail -2 gpurun_out/pytest_l.log
for v in 0 9 11 12 13; do
  timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-configs --variant $v > gpurun_out/bench_v$v.json 2> gpurun_out/bench_v$v.err
  python - $v <<'PY'
import json, sys
try:
    d = json.load(open(f"gpurun_out/bench_v{sys.argv[1]}.json"))
    print("variant", sys.argv[1], round(d["value"]), "evals/s", {k: round(x, 3) for k, x in d["stage_ms"].items()})
except Exception as e:
    print("FAILED", sys.argv[1], e)
PY
done
