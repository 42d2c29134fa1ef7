#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x -k "hist or skip or flags or search_matches or frames" > gpurun_out/pytest_hist.log 2>&1; tail -2 gpurun_out/pytest_hist.log
for p in 1 0; do for f in sky constant; do
  NMI_SKIP_PERSISTENT=$p timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-configs --frame $f 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('persistent $p', '$f', round(d['value']), {k: round(x, 3) for k, x in d['stage_ms'].items()})"
done; done
