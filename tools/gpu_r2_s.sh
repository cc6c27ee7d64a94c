#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_guard.py -m gpu -q -x > gpurun_out/r2s_pytest.log 2>&1; echo "pytest rc=$?"
grep -E "^(FAILED|ERROR)|^E  |passed|failed" gpurun_out/r2s_pytest.log | tail -12
