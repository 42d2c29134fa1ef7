#!/bin/bash
# round 2, run O: sub-binned records (one record per splat): parity + step time, tile_resolve at 16 vs 12 CTAs per SM
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x -k "parity or search or render" 2>&1 | tail -2
for lib in default tile12; do
  if [ $lib = default ]; then unset NMI_B200_LIB; else export NMI_B200_LIB=$PWD/orbslam2_nmi_b200/_lib/variants/$lib.so; fi
  python bench.py --no-configs --steps 20 --warmup 5 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$lib', d['value'], d['ms_per_step'], d['stage_ms'])"
done
