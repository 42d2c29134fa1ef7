#!/bin/bash
# round 2, run B: hist kernel capped at 64 registers -> does the two-context pipeline overlap now?  + NPP probe + ubench
mkdir -p gpurun_out
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
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_b.json 2> gpurun_out/bench_b.err; q gpurun_out/bench_b.json
timeout 300 python tools/exp_pipeline.py 20 > gpurun_out/exp_pipeline_b.txt 2>&1; tail -3 gpurun_out/exp_pipeline_b.txt
timeout 600 python tools/npp_warp_probe.py > gpurun_out/npp_probe.log 2>&1; tail -4 gpurun_out/npp_probe.log
timeout 120 ./tools/ubench_atoms > gpurun_out/ubench_atoms.txt 2>&1; cat gpurun_out/ubench_atoms.txt
