#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_api.py -m gpu -q -x -k "stiffness or sink" > gpurun_out/r2o_pytest.log 2>&1; echo "pytest rc=$?"
grep -E "^(FAILED|ERROR)|^E  |passed|failed" gpurun_out/r2o_pytest.log | tail -12
