#!/bin/bash
# round 2, run I (2 GPUs): bench arm under torchrun with the configs block, C++ NCCL host on 2 GPUs, compat tests
mkdir -p gpurun_out
nvidia-smi -L
timeout 600 python -m pytest tests/test_nccl_host_cpp.py tests/test_compat_cpp.py tests/test_reference_geometry.py -m gpu -q > gpurun_out/pytest_i.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_i.log; tail -4 gpurun_out/pytest_i.log
g++ -std=c++17 -O1 -I include -I /usr/local/cuda/include tests/cpp/test_nccl_host.cpp -L orbslam2_nmi_b200/_lib -lnmi_b200 -Wl,-rpath,$PWD/orbslam2_nmi_b200/_lib -L/usr/local/cuda/lib64 -lcudart -lnccl -lpthread -o /tmp/test_nccl_host && /tmp/test_nccl_host 0 2>&1 | tail -3
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err ) 2>&1 | grep real
tail -c 600 gpurun_out/bench_2gpu.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_2gpu.json"))
print(d["n_gpus"], "GPUs", round(d["value"]), "evals/s e2e", round(d["e2e"]["value"]), {k: round(x, 3) for k, x in d["stage_ms"].items()})
for k, v in d.get("configs", {}).items():
    print(k, json.dumps(v)[:600])
PY
