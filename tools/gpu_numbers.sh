#!/bin/bash
# every headline number of the round (no test suite: tools/gpu_final.sh runs that too)
mkdir -p gpurun_out
python __graft_entry__.py smoke 2>&1 | tail -1
python bench.py > gpurun_out/bench_default.log 2>&1; echo "bench default rc=$?"
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.log 2>&1; echo "bench ref rc=$?"
for wk in c2 c4 c5; do python bench.py --workload $wk --steps 3 --warmup 3 --cpu-seconds 6 > gpurun_out/bench_$wk.log 2>&1; echo "$wk rc=$?"; done
for wk in c3 c2 c4; do python bench.py --workload $wk --arithmetic fast --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${wk}_fast.log 2>&1; echo "$wk fast rc=$?"; done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/bench_*.log')):
    try: d=json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e: print(f,'?',e); continue
    r=d.get('roofline',{}); e=d.get('e2e') or {}; c=d.get('cpu_baseline') or {}
    print(f.split('/')[-1], 'value %.4g'%d['value'], 'ms %.2f'%d['ms_per_step'], 'e2e %.4g'%e.get('value',0), 'frac', round(r.get('frac',0) or 0,4), 'hbm', round((r.get('hbm') or {}).get('frac',0),4), 'cpu %.4g'%c.get('value',0), c.get('cores'))
PY
