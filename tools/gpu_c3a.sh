#!/bin/bash
mkdir -p gpurun_out
python tools/exp_c3_profile.py 6 > gpurun_out/c3_plain.log 2>&1; tail -2 gpurun_out/c3_plain.log
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 60 --csv --log-file gpurun_out/c3_launches.csv python tools/exp_c3_profile.py 4 > /dev/null 2>&1
ncu --set full --import-source on --clock-control none --profile-from-start off -k regex:"mesh_raster|mesh_shade|mesh_cull" -c 5 -f -o gpurun_out/prof_c3 python tools/exp_c3_profile.py 4 > gpurun_out/prof_ncu_c3.log 2>&1
tail -1 gpurun_out/prof_ncu_c3.log
