#!/bin/bash
# ncu --set full capture (with SASS source) of one run-kernel launch: gpurun -- 'bash tools/gpu_prof_one.sh <tag> <kernel regex> <bench args>'
# -> gpurun_out/prof_<tag>.ncu-rep (read with tools/stall_table.py / tools/ncu_summary.py)
mkdir -p gpurun_out
TAG=$1; KRE=$2; shift 2
CMDP="python bench.py $* --steps 1 --warmup 1 --no-e2e --no-cpu-baseline"
$CMDP > gpurun_out/plain_$TAG.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:$KRE -s 1 -c 1 -o gpurun_out/prof_$TAG -f $CMDP > gpurun_out/ncu_$TAG.log 2>&1
echo "$TAG rc=$?"
