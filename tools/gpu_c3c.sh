#!/bin/bash
for z in 8 16 32 48; do echo "zbuf MB $z"; NMI_ZBUF_MB=$z python tools/exp_c3_profile.py 6 2>&1 | tail -2 | head -1; done
