#!/bin/bash
# round 2, run J: tile_resolve 32-bit filter: parity (render tests) + bench + ncu of the render-stage kernels
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_compat_cpp.py -m gpu -q -x > gpurun_out/pytest_j.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_j.log; tail -3 gpurun_out/pytest_j.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-configs > gpurun_out/bench_j.json 2> gpurun_out/bench_j.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_j.json"))
print(round(d["value"]), "evals/s e2e", round(d["e2e"]["value"]), {k: round(x, 3) for k, x in d["stage_ms"].items()})
PY
bash tools/gpu_prof_render.sh
