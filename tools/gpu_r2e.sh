#!/bin/bash
# round 2, run E: the new bench line (configs block, roofline) + GPU tests touched by the refactor
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_hist_fullsize.py -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
( time timeout 900 python bench.py > gpurun_out/bench_e.json 2> gpurun_out/bench_e.err ) 2>&1 | grep real
tail -c 400 gpurun_out/bench_e.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_e.json"))
print(round(d["value"]), "evals/s e2e", round(d["e2e"]["value"]), "pageable", round(d["e2e"]["pageable_frame"]["value"]), {k: round(x, 3) for k, x in d["stage_ms"].items()})
r = d["roofline"]; print("roofline", r["bound"], round(r["achieved"],1), "/", round(r["peak"],1), r["unit"], "frac", round(r["frac"],3), "| hbm frac", round(r["hbm"]["frac"],3), r["peak_detail"]["source"])
for k, v in d.get("configs", {}).items():
    print(k, json.dumps(v)[:700])
print("cpu", json.dumps(d.get("cpu_baseline", {}))[:400])
PY
