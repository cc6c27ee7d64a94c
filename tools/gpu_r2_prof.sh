#!/bin/bash
# round 2 profile artefacts with the current kernels: launch list of the default bench + ncu --set full of every config
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-fast"
$CMD > gpurun_out/plain_c3.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/r02_launches_c3.csv $CMD > gpurun_out/ncu_l.log 2>&1
echo "launch list rc=$?"
prof() {  # name, kernel regex, bench args
  local CMDP="python bench.py $3 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-fast"
  $CMDP > gpurun_out/plain_$1.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:$2 -s 1 -c 1 -o gpurun_out/prof_$1 -f $CMDP > gpurun_out/ncu_$1.log 2>&1
  echo "$1 rc=$?"
}
prof r2_c5rec ni_k_large_run "--workload c5 --traj 1184 --record"
prof r2_c5 ni_k_large_run "--workload c5 --traj 2048"
prof r2_c3 ni_k_run "--traj 2000000"
prof r2_c3fast ni_k_run "--traj 2000000 --arithmetic fast"
prof r2_c4 ni_k_run "--workload c4 --traj 400000"
prof r2_c2 ni_k_run "--workload c2 --traj 300000"
ls -la gpurun_out/prof_r2_* | awk '{print $5, $9}'
