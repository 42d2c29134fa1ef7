#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "mesh" > gpurun_out/pytest_mesh.log 2>&1; tail -5 gpurun_out/pytest_mesh.log
