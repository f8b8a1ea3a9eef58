#!/bin/bash
# Round 2, call D (1 GPU): refactored blocked kernels + persistent solver: suite, microbench, solve timings.
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --durations=8 > gpurun_out/r2d_pytest_gpu.txt 2>&1
echo "pytest rc=$?" >> gpurun_out/r2d_pytest_gpu.txt
tail -30 gpurun_out/r2d_pytest_gpu.txt
timeout 300 python scripts/microbench.py 2>&1 | tail -4 > gpurun_out/r2d_microbench.txt; cat gpurun_out/r2d_microbench.txt
timeout 300 python scripts/time_solve.py > gpurun_out/r2d_time_solve.txt 2>&1; cat gpurun_out/r2d_time_solve.txt
