#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/pytest_gpu.log | tail -25
python tools/parity_report.py > gpurun_out/parity.txt 2>&1; echo "parity rc=$?"
grep -E "heat|burgers|c3_lorenz_t1|default_tol|readme" gpurun_out/parity.txt
timeout 600 python bench.py --workload c5 --steps 3 --warmup 3 --cpu-seconds 6 > gpurun_out/bench_c5.log 2>&1; echo "bench c5 rc=$?"; tail -2 gpurun_out/bench_c5.log
