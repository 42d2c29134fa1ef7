#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log; tail -3 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; tail -c 300 gpurun_out/bench_default.err
python -c "
import json
d = json.load(open('gpurun_out/bench_default.json'))
print(round(d['value']), 'e2e', round(d['e2e']['value']), 'C5', d['configs']['C5']['ms_per_frame_mean'], d['configs']['C5']['ms_per_frame_max'], 'C3', d['configs']['C3']['search_ms'], d['configs']['C3']['parity'])"
