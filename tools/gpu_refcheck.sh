#!/bin/bash
# the reference's own kernels (oracle/_ref) against our CUDA path and the oracle, then the usual suite + both bench arms
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_reference_kernels.py -m gpu -q > gpurun_out/pytest_refkernels.log 2>&1; echo "exit $?" >> gpurun_out/pytest_refkernels.log
tail -25 gpurun_out/pytest_refkernels.log
timeout 1200 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_reference_kernels.py > gpurun_out/pytest_gpu.log 2>&1; echo "exit $?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
timeout 600 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; tail -c 1500 gpurun_out/bench_default.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; tail -c 900 gpurun_out/bench_reference.json
