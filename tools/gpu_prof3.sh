#!/bin/bash
mkdir -p gpurun_out
CMD3="python bench.py --workload c5 --traj 2048 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline"
$CMD3 > gpurun_out/plain3.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:ni_k_large_run -s 1 -c 1 -o gpurun_out/prof_c5 -f $CMD3 > gpurun_out/ncu3.log 2>&1
echo "c5 full rc=$?"
