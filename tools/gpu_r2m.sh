#!/bin/bash
# round 2, run M: single-evaluation cluster kernel: parity (every test that evaluates single pairs) + C1 latency A/B
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_m.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_m.log; tail -6 gpurun_out/pytest_m.log
for cl in 1 0; do
NMI_EVAL_CLUSTER=$cl python - <<'PY'
import os, sys, time, json
sys.path.insert(0, ".")
import bench
r = bench.config_c1(0)
print("NMI_EVAL_CLUSTER", os.environ["NMI_EVAL_CLUSTER"], "eval_pair_call_us", round(r["eval_pair_call_us"], 1), "search wall ms", round(r["search_ms_wall"], 4), "device", round(r["search_ms_device"], 4), r["stage_ms"]["hist_score"], r["parity"])
PY
done
