#!/bin/bash
# round-1 profile artefacts: launch list of the default bench command + full captures of the two run kernels
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/plain_c3.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_c3.csv $CMD > gpurun_out/ncu_l.log 2>&1
echo "launch list rc=$?"
CMD2="python bench.py --traj 2000000 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline"
$CMD2 > gpurun_out/plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:ni_k_run -s 1 -c 1 -o gpurun_out/prof_c3 -f $CMD2 > gpurun_out/ncu2.log 2>&1
echo "c3 full rc=$?"
CMD3="python bench.py --workload c5 --traj 2048 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline"
$CMD3 > gpurun_out/plain3.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:ni_k_large_run -s 1 -c 1 -o gpurun_out/prof_c5 -f $CMD3 > gpurun_out/ncu3.log 2>&1
echo "c5 full rc=$?"
tail -1 gpurun_out/plain3.log
