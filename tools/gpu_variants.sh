#!/bin/bash
# A/B of library variants built by tools/build_variant.py: bench stage times for each
mkdir -p gpurun_out
for v in default $(ls orbslam2_nmi_b200/_lib/variants/ | sed 's/\.so$//'); do
  if [ "$v" = default ]; then unset NMI_B200_LIB; else export NMI_B200_LIB=$PWD/orbslam2_nmi_b200/_lib/variants/$v.so; fi
  timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-configs > gpurun_out/bench_var_$v.json 2> gpurun_out/bench_var_$v.err
  python - $v <<'PY'
import json, sys
try:
    d = json.load(open(f"gpurun_out/bench_var_{sys.argv[1]}.json"))
    print(sys.argv[1], round(d["value"]), "evals/s", {k: round(x, 3) for k, x in d["stage_ms"].items()})
except Exception as e:
    print("FAILED", sys.argv[1], e)
PY
done
