#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
grep -E "^(FAILED|ERROR)|^E  |passed|failed" gpurun_out/pytest_gpu.log | tail -12
for wk in c5; do
timeout 900 python bench.py --workload $wk --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$wk.log 2>&1; echo "bench $wk rc=$?"; python - <<PY
import json
d=json.loads(open('gpurun_out/bench_$wk.log').read().strip().splitlines()[-1])
print('$wk', 'Msteps/s', d['value']/1e6, 'e2e', d['e2e']['value']/1e6, 'kernel_ms', d['roofline']['kernel_ms'], 'frac', d['roofline']['frac'])
PY
done
