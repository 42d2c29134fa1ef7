#!/bin/bash
# round 2, run C: same L1/shared carve-out on every kernel -> do render and histogram CTAs share SMs now?
mkdir -p gpurun_out
for co in 1 0; do
  echo "== NMI_CARVEOUT=$co"
  NMI_CARVEOUT=$co timeout 300 python tools/exp_pipeline.py 20 2>&1 | tail -2
done
