#!/bin/bash
# first contact with the GPU: smoke, parity table, tests, a small bench
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"
tail -3 gpurun_out/smoke.log
python tools/parity_report.py --oracle-n 20000 > gpurun_out/parity.txt 2>&1; echo "parity rc=$?"
cat gpurun_out/parity.txt
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/pytest_gpu.log
python bench.py --traj 1000000 --steps 3 --warmup 3 > gpurun_out/bench_1m.log 2>&1; echo "bench rc=$?"
tail -3 gpurun_out/bench_1m.log
