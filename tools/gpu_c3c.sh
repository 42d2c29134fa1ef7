#!/bin/bash
python tools/exp_c3_profile.py 6 2>&1 | tail -2 | head -1
for n in 4 2; do echo "b64 copies $n"; NMI_B200_LIB=orbslam2_nmi_b200/_lib/variants/b64c$n.so python tools/exp_c3_profile.py 6 2>&1 | tail -2 | head -1; done
