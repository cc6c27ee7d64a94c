#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
grep -E "^(FAILED|ERROR)|^E  |passed|failed" gpurun_out/pytest_gpu.log | tail -12
for wk in c3 c2 c4; do for ar in strict fast; do
python bench.py --workload $wk --arithmetic $ar --steps 3 --warmup 3 --no-e2e --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.readlines()[-1]); print('$wk $ar', round(d['value']/1e9,2), round(d['roofline']['kernel_ms'],2), round(d['roofline']['frac'],4))"
done; done
