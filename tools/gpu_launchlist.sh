#!/bin/bash
mkdir -p gpurun_out
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_short.json 2> gpurun_out/bench_short.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
python tools/summarize_launches.py gpurun_out/launches.csv gpurun_out/launches_summary.txt "bench.py --steps 2 --warmup 3"
