#!/bin/bash
for v in 1 2 3 4; do echo "shade variant $v"; NMI_SHADE_V=$v python tools/exp_c3_profile.py 6 2>&1 | tail -2 | head -1; done
