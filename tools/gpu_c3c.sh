#!/bin/bash
mkdir -p gpurun_out
NMI_SHADE_QUAD=1 python tools/exp_c3_profile.py 6 2>&1 | tail -2
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 60 --csv --log-file gpurun_out/c3_launches.csv python tools/exp_c3_profile.py 4 > /dev/null 2>&1
