#!/bin/bash
# round 2, run K: histogram hot loop with the cheaper address arithmetic (8 instead of 10.5 issue slots per pixel)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_hist_fullsize.py tests/test_gpu_reference_kernels.py -m gpu -q -x > gpurun_out/pytest_k.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_k.log; tail -3 gpurun_out/pytest_k.log
for f in textured sky; do
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-configs --frame $f > gpurun_out/bench_k_$f.json 2> gpurun_out/bench_k_$f.err
python - $f <<'PY'
import json, sys
d = json.load(open(f"gpurun_out/bench_k_{sys.argv[1]}.json"))
print(sys.argv[1], round(d["value"]), "evals/s e2e", round(d["e2e"]["value"]), {k: round(x, 3) for k, x in d["stage_ms"].items()})
PY
done
./orbslam2_nmi_b200/_lib/ubench_atoms 2>&1 | tail -5
