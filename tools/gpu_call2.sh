#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
for v in 0 4 5 6 2 7; do
  timeout 300 python bench.py --steps 5 --warmup 3 --variant $v --no-cpu-baseline > gpurun_out/bench_v${v}_textured.json 2> gpurun_out/bench_v${v}_textured.err
  python - <<PY
import json
d=json.load(open("gpurun_out/bench_v${v}_textured.json"))
print("variant $v textured", round(d["value"]), "evals/s", {k: round(x,3) for k,x in d["stage_ms"].items()})
PY
done
for f in uniform constant; do
 for v in 4 6; do
  timeout 300 python bench.py --steps 3 --warmup 3 --variant $v --frame $f --no-cpu-baseline > gpurun_out/bench_v${v}_$f.json 2> gpurun_out/bench_v${v}_$f.err
  python - <<PY
import json
d=json.load(open("gpurun_out/bench_v${v}_$f.json"))
print("variant $v $f", round(d["value"]), "evals/s", {k: round(x,3) for k,x in d["stage_ms"].items()})
PY
 done
done
V=${PROF_VARIANT:-0}
python tools/profile_run.py $V > gpurun_out/prof_plain.log 2>&1 && \
ncu --set full --import-source on --clock-control none -k regex:"joint_hist|project_splat|resolve" -s 3 -c 3 -f -o gpurun_out/prof_r1_v$V python tools/profile_run.py $V > gpurun_out/prof_ncu.log 2>&1
tail -3 gpurun_out/prof_plain.log; tail -3 gpurun_out/prof_ncu.log
