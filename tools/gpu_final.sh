#!/bin/bash
# round-end evidence run: tests, smoke, bench (ours + reference arm), ncu launch list + full captures
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; tail -c 2600 gpurun_out/bench_default.json
# launch list of the same command line (only after it exited 0 without ncu)
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_short.json 2> gpurun_out/bench_short.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
# full captures of one steady-state search (profile_run.py brackets its last search with cudaProfilerStart/Stop)
python tools/profile_run.py 0 > gpurun_out/prof_plain.log 2>&1 && \
ncu --set full --import-source on --clock-control none --profile-from-start off -k regex:"cull_|bin_kernel|tile_resolve|warp_kernel|argmax" -c 12 -f -o gpurun_out/prof_render python tools/profile_run.py 0 > gpurun_out/prof_ncu_render.log 2>&1 && \
ncu --set full --import-source on --clock-control none --profile-from-start off -k regex:"joint_hist" -c 1 -f -o gpurun_out/prof_hist python tools/profile_run.py 0 > gpurun_out/prof_ncu_hist.log 2>&1
tail -1 gpurun_out/prof_ncu_render.log; tail -1 gpurun_out/prof_ncu_hist.log
# histogram contention cases and the other BASELINE configs, for the record
CASES="textured_40_0 textured_40_1 sky_40_0 sky_40_1 sky_12_0 sky_12_1 constant_40_0 constant_40_1" bash tools/gpu_stress.sh > gpurun_out/hist_stress.txt 2>&1
python tools/run_configs.py > gpurun_out/configs.jsonl 2> gpurun_out/configs.err
