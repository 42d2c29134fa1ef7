#!/bin/bash
# ncu captures of the render-stage kernels and of the histogram kernel (second search)
mkdir -p gpurun_out
V=${PROF_VARIANT:-0}
python tools/profile_run.py $V > gpurun_out/prof_plain.log 2>&1 && \
ncu --set full --import-source on --clock-control none -k regex:"cull_|bin_kernel|tile_resolve|warp_kernel" -s ${SKIP:-36} -c 8 -f -o gpurun_out/prof_render python tools/profile_run.py $V > gpurun_out/prof_ncu_render.log 2>&1 && \
ncu --set full --import-source on --clock-control none -k regex:"joint_hist" -s 1 -c 1 -f -o gpurun_out/prof_hist python tools/profile_run.py $V > gpurun_out/prof_ncu_hist.log 2>&1
tail -2 gpurun_out/prof_plain.log | cut -c1-300; tail -2 gpurun_out/prof_ncu_render.log; tail -2 gpurun_out/prof_ncu_hist.log
for v in ${VARIANTS:-1}; do
  timeout 300 python bench.py --steps 5 --warmup 3 --variant $v --no-cpu-baseline > gpurun_out/bench_v${v}_textured.json 2> gpurun_out/bench_v${v}_textured.err
  python - <<PY
import json
d=json.load(open("gpurun_out/bench_v${v}_textured.json"))
print("variant $v textured", round(d["value"]), "evals/s", {k: round(x,3) for k,x in d["stage_ms"].items()})
PY
done
