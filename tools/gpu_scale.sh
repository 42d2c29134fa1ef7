#!/bin/bash
# weak-scaling check: bench at N = 4 and 8 (run under `gpurun --gpus 8`), as the driver launches it
mkdir -p gpurun_out
for n in ${NS:-4 8}; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + n)) \
    bench.py --gpus $n --steps 10 --warmup 3 > gpurun_out/bench_${n}gpu.json 2> gpurun_out/bench_${n}gpu.err
  echo "N=$n rc=$?"; tail -c 700 gpurun_out/bench_${n}gpu.json | head -c 700; echo
done
