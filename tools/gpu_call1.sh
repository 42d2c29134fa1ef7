#!/bin/bash
# first GPU call: parity tests, atomics microbenchmark, histogram-variant sweep, launch list
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 120 ./tools/ubench_atoms > gpurun_out/ubench_atoms.txt 2>&1; tail -20 gpurun_out/ubench_atoms.txt
for v in 0 1 2 3; do
  timeout 300 python bench.py --steps 5 --warmup 3 --variant $v --no-cpu-baseline > gpurun_out/bench_v${v}_textured.json 2> gpurun_out/bench_v${v}_textured.err
  tail -c 1500 gpurun_out/bench_v${v}_textured.json
done
for f in uniform constant; do
  timeout 300 python bench.py --steps 3 --warmup 3 --variant 0 --frame $f --no-cpu-baseline > gpurun_out/bench_v0_$f.json 2> gpurun_out/bench_v0_$f.err
  tail -c 600 gpurun_out/bench_v0_$f.json
done
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
tail -c 2500 gpurun_out/bench_default.json
timeout 300 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_reference.json 2>&1
