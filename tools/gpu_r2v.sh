#!/bin/bash
mkdir -p gpurun_out
for i in 1 2; do
python bench.py --no-cpu-baseline > gpurun_out/bench_nc.json 2> gpurun_out/bench_nc.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_nc.json"))
c = d["configs"]["C5"]
print(round(d["value"]), "C5 mean", round(c["ms_per_frame_mean"], 3), "median", round(c["ms_per_frame_median"], 3), "max", round(c["ms_per_frame_max"], 2), c["slowest_frames"], "small", round(d["configs"]["C5_small_grid"]["ms_per_frame_mean"], 3))
PY
done
