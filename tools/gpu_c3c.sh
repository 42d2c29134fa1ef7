#!/bin/bash
python tools/exp_c3_profile.py 6 2>&1 | tail -2 | head -1
for m in 5 6 8; do echo "raster minb $m"; NMI_B200_LIB=orbslam2_nmi_b200/_lib/variants/rast$m.so python tools/exp_c3_profile.py 6 2>&1 | tail -2 | head -1; done
