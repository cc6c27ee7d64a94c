#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/pytest_gpu.log | tail -8
for wk in c2 c4; do
timeout 900 python bench.py --workload $wk --steps 3 --warmup 3 --cpu-seconds 6 > gpurun_out/bench_$wk.log 2>&1; echo "bench $wk rc=$?"; tail -1 gpurun_out/bench_$wk.log
done
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.log 2>&1; echo "ref rc=$?"; tail -1 gpurun_out/bench_ref.log
nproc; lscpu | grep "Model name"
