#!/bin/bash
# round 2, call X: executed warp-instructions of unordered against cost-ordered launches of the C3 run kernel (ncu)
mkdir -p gpurun_out
ncu --metrics smsp__inst_executed.sum,gpu__time_duration.sum --clock-control none -k regex:ni_k_run --csv --log-file gpurun_out/r02_ordered_inst.csv \
  python bench.py --traj 2000000 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-fast > gpurun_out/q.log 2>&1; echo "rc=$?"
grep -c ni_k_run gpurun_out/r02_ordered_inst.csv
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r02_ordered_inst.csv')) if len(r)>5]
h=rows[0]; ci={n:i for i,n in enumerate(h)}
out={}
for r in rows[1:]:
    out.setdefault(r[ci['ID']],{})[r[ci['Metric Name']]]=r[ci['Metric Value']]
for k,v in out.items(): print(k, v)
PY
