#!/bin/bash
mkdir -p gpurun_out
python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; tail -c 300 gpurun_out/bench_default.err
python -c "
import json
d = json.load(open('gpurun_out/bench_default.json'))
print(round(d['value']), 'e2e', round(d['e2e']['value']), 'C5', d['configs']['C5']['ms_per_frame_mean'], d['configs']['C5']['ms_per_frame_max'], 'C3', d['configs']['C3']['search_ms'])"
