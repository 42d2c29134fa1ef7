#!/bin/bash
# full ncu capture of one steady-state search's render-stage kernels (profile_run.py brackets the last search)
mkdir -p gpurun_out
python tools/profile_run.py 0 > gpurun_out/prof_plain.log 2>&1 && \
ncu --set full --import-source on --clock-control none --profile-from-start off -k regex:"cull_|bin_kernel|tile_resolve|warp_kernel|argmax|image_mode" -c 12 -f -o gpurun_out/prof_render python tools/profile_run.py 0 > gpurun_out/prof_ncu_render.log 2>&1
tail -2 gpurun_out/prof_plain.log; tail -1 gpurun_out/prof_ncu_render.log
