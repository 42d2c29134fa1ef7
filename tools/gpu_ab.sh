#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; tail -2 gpurun_out/pytest_gpu.log
run() { # name, env
  timeout 300 env $2 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$1.json 2> gpurun_out/bench_$1.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_$1.json"))
    print("$1", round(d["value"]), "evals/s", {k: round(x,3) for k,x in d["stage_ms"].items()})
except Exception as e:
    print("$1 FAILED", e, open("gpurun_out/bench_$1.err").read()[-300:])
PY
}
run base "X=1"
run exp "NMI_B200_LIB=$PWD/orbslam2_nmi_b200/_lib_exp/libnmi_b200.so"
run base2 "X=1"
