#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --durations=8 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
grep -E "^(FAILED|ERROR)|^E  |passed|failed|s call" gpurun_out/pytest_gpu.log | tail -24
python bench.py --steps 3 --warmup 3 > gpurun_out/bench_c3.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/bench_c3.log | cut -c1-1800
