#!/bin/bash
# round 2, call N: in-kernel terminal sink + record rows of the large kernels: full GPU suite
mkdir -p gpurun_out
timeout 1700 python -m pytest tests -m gpu -q > gpurun_out/r2n_pytest_gpu.log 2>&1; echo "pytest rc=$?"
grep -E "^(FAILED|ERROR)|^E  |passed|failed" gpurun_out/r2n_pytest_gpu.log | tail -12
