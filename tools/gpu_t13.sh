#!/bin/bash
mkdir -p gpurun_out
for wk in c3 c5 c2; do
timeout 900 python bench.py --workload $wk --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$wk.log 2>&1; echo "bench $wk rc=$?"; python - <<PY
import json
d=json.loads(open('gpurun_out/bench_$wk.log').read().strip().splitlines()[-1])
print('$wk', 'value', round(d['value']/1e9,3), 'e2e', round(d['e2e']['value']/1e9,3), 'ms', round(d['ms_per_step'],2), 'e2e ms', round(d['e2e']['ms_per_step'],2))
PY
done
python -m pytest tests/test_gpu_api.py -q -k structref 2>&1 | tail -1
