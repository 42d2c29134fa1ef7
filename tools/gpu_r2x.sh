#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_nc.json 2> gpurun_out/bench_nc.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_nc.json"))
print(round(d["value"]), "evals/s e2e", round(d["e2e"]["value"]), "pageable", round(d["e2e"]["pageable_frame"]["value"]), {k: round(x, 3) for k, x in d["stage_ms"].items()})
c = d["configs"]
print("C5", round(c["C5"]["ms_per_frame_mean"], 3), "small", round(c["C5_small_grid"]["ms_per_frame_mean"], 3), c["C5_small_grid"]["per_level_rank0_us"][0], "C3", round(c["C3"]["search_ms"], 3), "C1", c["C1"]["search_ms_wall"], c["C1"]["eval_pair_call_us"])
PY
