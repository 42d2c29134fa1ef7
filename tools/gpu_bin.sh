#!/bin/bash
for lib in "" orbslam2_nmi_b200/_lib/variants/bin128.so orbslam2_nmi_b200/_lib/variants/bin512.so; do
  echo "lib ${lib:-default}"
  env ${lib:+NMI_B200_LIB=$lib} python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-configs 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(round(d['value']), {k: round(x, 3) for k, x in d['stage_ms'].items()})"
done
