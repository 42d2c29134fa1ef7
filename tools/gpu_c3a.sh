#!/bin/bash
mkdir -p gpurun_out
ncu --set full --import-source on --clock-control none --profile-from-start off -k regex:"mesh_shade" -c 1 -f -o gpurun_out/prof_c3_quad python tools/exp_c3_profile.py 4 > gpurun_out/prof_ncu_c3.log 2>&1
tail -1 gpurun_out/prof_ncu_c3.log
