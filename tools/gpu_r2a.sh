#!/bin/bash
# round 2, run A: parity suite on the new render stage, quick bench, block-cull A/B, two-context pipeline
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
q() {
python - "$1" <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1]))
    print(sys.argv[1], round(d["value"]), "evals/s e2e", round(d["e2e"]["value"]), {k: round(x, 3) for k, x in d["stage_ms"].items()}, "launches", d["gpu_launches"])
except Exception as e:
    print("FAILED", sys.argv[1], e)
PY
}
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_a.json 2> gpurun_out/bench_a.err; q gpurun_out/bench_a.json
NMI_BLOCK_CULL=0 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_a_noblock.json 2> gpurun_out/bench_a_noblock.err; q gpurun_out/bench_a_noblock.json
timeout 300 python tools/exp_pipeline.py 20 > gpurun_out/exp_pipeline.txt 2>&1; tail -3 gpurun_out/exp_pipeline.txt
