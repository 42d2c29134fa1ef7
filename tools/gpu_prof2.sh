#!/bin/bash
mkdir -p gpurun_out
python tools/profile_run.py 0 3 > gpurun_out/prof_plain.log 2>&1 && \
ncu --set full --import-source on --clock-control none -k regex:"joint_hist" -s 2 -c 1 -f -o gpurun_out/prof_hist python tools/profile_run.py 0 3 > gpurun_out/prof_ncu_hist.log 2>&1 && \
ncu --set full --import-source on --clock-control none -k regex:"cull_|bin_kernel|tile_resolve|warp_kernel" -s ${SKIP:-44} -c 6 -f -o gpurun_out/prof_render python tools/profile_run.py 0 3 > gpurun_out/prof_ncu_render.log 2>&1
tail -1 gpurun_out/prof_plain.log | cut -c1-300; tail -1 gpurun_out/prof_ncu_hist.log; tail -1 gpurun_out/prof_ncu_render.log
