#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_reference_kernels.py -m gpu -q > gpurun_out/pytest_refkernels.log 2>&1; echo "exit $?" >> gpurun_out/pytest_refkernels.log
tail -25 gpurun_out/pytest_refkernels.log
