#!/bin/bash
# tests + parity table + bench + ncu capture of the run kernel on a small C3 launch
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/pytest_gpu.log | tail -25
python tools/parity_report.py --oracle-n 20000 > gpurun_out/parity.txt 2>&1; echo "parity rc=$?"
grep -E "c2_|c3_|c4_|readme|lorenz_single" gpurun_out/parity.txt
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_c3.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/bench_c3.log
CMD="python bench.py --traj 2000000 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline"
$CMD > gpurun_out/plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:ni_k_run -s 1 -c 1 -o gpurun_out/prof_run -f $CMD > gpurun_out/ncu.log 2>&1
echo "ncu rc=$?"; tail -2 gpurun_out/ncu.log
