#!/bin/bash
# A/B of histogram kernel variants on the bench workload + the GPU suite
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "exit $?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
for v in ${VARIANTS:-0 8}; do
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --variant $v > gpurun_out/bench_v$v.json 2> gpurun_out/bench_v$v.err
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_v$v.json"))
    print("variant $v", round(d["value"]), "evals/s e2e", round(d["e2e"]["value"]), {k: round(x,3) for k,x in d["stage_ms"].items()}, d["config"]["winner"])
except Exception as e:
    print("FAILED", e, open("gpurun_out/bench_v$v.err").read()[-600:])
PY
done
for v in 0 8; do NMI_EXP_VARIANT=$v timeout 120 python tools/exp_hist_overhead.py tiny C1 2>&1 | tail -2; done
