#!/bin/bash
for cfg in "2 64" "3 64" "4 64" "3 96" "4 128" "4 96"; do set -- $cfg; echo "streams $1 zbuf $2"; NMI_MESH_STREAMS=$1 NMI_ZBUF_MB=$2 python tools/exp_c3_profile.py 6 2>&1 | tail -2 | head -1; done
