#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
for f in textured uniform constant; do
 for v in ${VARIANTS:-0 5 4 6}; do
  timeout 300 python bench.py --steps 5 --warmup 3 --variant $v --frame $f --no-cpu-baseline > gpurun_out/bench_v${v}_$f.json 2> gpurun_out/bench_v${v}_$f.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_v${v}_$f.json"))
    print("variant $v $f", round(d["value"]), "evals/s e2e", round(d["e2e"]["value"]), {k: round(x,3) for k,x in d["stage_ms"].items()})
except Exception as e:
    print("variant $v $f FAILED", e, open("gpurun_out/bench_v${v}_$f.err").read()[-400:])
PY
 done
done
for V in ${PROF_VARIANTS:-0 5}; do
python tools/profile_run.py $V > gpurun_out/prof_plain_v$V.log 2>&1 && \
ncu --set full --import-source on --clock-control none -k regex:"joint_hist|project_splat|resolve|warp_kernel|cull" -s 5 -c 5 -f -o gpurun_out/prof_v$V python tools/profile_run.py $V > gpurun_out/prof_ncu_v$V.log 2>&1
tail -2 gpurun_out/prof_plain_v$V.log | cut -c1-300; tail -2 gpurun_out/prof_ncu_v$V.log
done
