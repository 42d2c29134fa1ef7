#!/bin/bash
# round 2, run H: textured mesh, Rendering<1> drop-in, C++ NCCL host, full suite, bench with the textured C3
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -30 gpurun_out/pytest_gpu.log
( time timeout 900 python bench.py --no-cpu-baseline > gpurun_out/bench_h.json 2> gpurun_out/bench_h.err ) 2>&1 | grep real
tail -c 300 gpurun_out/bench_h.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_h.json"))
print(round(d["value"]), "evals/s e2e", round(d["e2e"]["value"]), {k: round(x, 3) for k, x in d["stage_ms"].items()})
for k in ("C3", "C1"):
    print(k, json.dumps(d.get("configs", {}).get(k))[:900])
PY
