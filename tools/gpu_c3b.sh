#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "mesh or Mesh or textured or compat" > gpurun_out/pytest_mesh.log 2>&1; tail -2 gpurun_out/pytest_mesh.log
python tools/exp_c3_profile.py 6 > gpurun_out/c3_plain.log 2>&1; tail -2 gpurun_out/c3_plain.log
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 120 --csv --log-file gpurun_out/c3_launches.csv python tools/exp_c3_profile.py 4 > /dev/null 2>&1
ncu --set full --import-source on --clock-control none --profile-from-start off -k regex:"mesh_vertices|mesh_raster|mesh_shade" -c 3 -f -o gpurun_out/prof_c3 python tools/exp_c3_profile.py 4 > gpurun_out/prof_ncu_c3.log 2>&1
