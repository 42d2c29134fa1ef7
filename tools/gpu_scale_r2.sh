#!/bin/bash
# round 2 multi-GPU evidence (run under `gpurun --gpus N`): the bench as the driver launches it (weak scaling of the C2
# step, C4 / C5 / C5 small grid sharded over the ranks inside the same line) and the C++ NCCL host test
mkdir -p gpurun_out
for n in ${NS:-2}; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + n)) \
    bench.py --gpus $n --steps 10 --warmup 3 > gpurun_out/bench_${n}gpu.json 2> gpurun_out/bench_${n}gpu.err
  echo "N=$n rc=$?"
  python - $n <<'PY'
import json, sys
n = sys.argv[1]
d = json.loads(open(f"gpurun_out/bench_{n}gpu.json").read().strip().splitlines()[-1])
print("N", n, "value", round(d["value"]), "ms/step", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"]))
for k, v in d.get("configs", {}).items():
    print(" ", k, {a: (round(b, 3) if isinstance(b, float) else b) for a, b in v.items() if a in ("search_ms", "evals_per_s", "ms_per_frame_mean", "ms_per_frame_median", "ms_per_frame_p99", "error")})
    if k == "C5_small_grid":
        print("   per level:", [{a: round(b) for a, b in lv.items()} for lv in v.get("per_level_rank0_us", [])])
PY
done
timeout 600 python -m pytest tests/test_nccl_host_cpp.py -m gpu -q 2>&1 | tail -2
