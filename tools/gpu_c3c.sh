#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "mesh or Mesh or textured or compat" > gpurun_out/pytest_mesh.log 2>&1; tail -2 gpurun_out/pytest_mesh.log
python tools/exp_c3_profile.py 6 2>&1 | tail -2
NMI_CULL_PASSES=3 python tools/exp_c3_profile.py 6 2>&1 | tail -2 | head -1
