#!/bin/bash
# round 2, run G: warp kernel without XU conversions: parity + NPP + bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
for wt in 1 0; do
NMI_WARP_TEX=$wt timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-configs > gpurun_out/bench_g$wt.json 2> gpurun_out/bench_g$wt.err
python - $wt <<'PY'
import json, sys
try:
    d = json.load(open(f"gpurun_out/bench_g{sys.argv[1]}.json"))
    print("warp_tex", sys.argv[1], round(d["value"]), "evals/s", {k: round(x, 3) for k, x in d["stage_ms"].items()})
except Exception as e:
    print("FAILED", e)
PY
done
