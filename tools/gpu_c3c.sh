#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "mesh or Mesh or textured or compat" > gpurun_out/pytest_mesh.log 2>&1; tail -2 gpurun_out/pytest_mesh.log
for l in 1 0; do echo "tv layout $l"; NMI_MESH_TV_LAYOUT=$l python tools/exp_c3_profile.py 6 2>&1 | tail -2 | head -1; done
