#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
grep -E "^(FAILED|ERROR)|^E  |passed|failed" gpurun_out/pytest_gpu.log | tail -24
for wk in c3 c2 c4; do for ar in strict fast; do
timeout 900 python bench.py --workload $wk --arithmetic $ar --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${wk}_$ar.log 2>&1; echo "bench $wk $ar rc=$?"; python - <<PY
import json
d=json.loads(open('gpurun_out/bench_${wk}_$ar.log').read().strip().splitlines()[-1])
print('$wk $ar', 'Gsteps/s', round(d['value']/1e9,2), 'e2e', round(d['e2e']['value']/1e9,2), 'kernel_ms', round(d['roofline']['kernel_ms'],2), 'frac', round(d['roofline']['frac'],4))
PY
done; done
