#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "flags or mesh or randomized or search_matches" > gpurun_out/pytest_b64.log 2>&1; tail -2 gpurun_out/pytest_b64.log
python tools/exp_c3_profile.py 6 2>&1 | tail -2 | head -1
