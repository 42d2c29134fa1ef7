#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x -k "mesh" > gpurun_out/pytest_mesh.log 2>&1; tail -2 gpurun_out/pytest_mesh.log
for mb in 64; do
NMI_ZBUF_MB=$mb python tools/run_configs.py --only C3 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: print(l[-300:]); continue
    print('zbuf_mb', $mb, round(d['search_ms'],3), 'ms', round(d['evals_per_s']), {k: round(v,3) for k,v in d['stage_ms'].items()})
"
done
