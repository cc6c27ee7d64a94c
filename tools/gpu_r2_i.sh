#!/bin/bash
# round 2, call I: full GPU suite + smoke + default bench on the state-setter / 16-byte pack commit
mkdir -p gpurun_out
timeout 1700 python -m pytest tests -m gpu -q -x > gpurun_out/r2i_pytest_gpu.log 2>&1; echo "pytest rc=$?"
grep -E "^(FAILED|ERROR)|^E  |passed|failed" gpurun_out/r2i_pytest_gpu.log | tail -12
python __graft_entry__.py smoke 2>&1 | tail -1
python bench.py > gpurun_out/r2i_bench.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/r2i_bench.log | cut -c1-1500
