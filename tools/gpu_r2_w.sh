#!/bin/bash
# round 2, call W: C5 end to end (host buffers) against its device time; terminal-record tests
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_api.py tests/test_gpu_large_features.py -m gpu -q -x -k "terminal or large or recording or auxiliary" > gpurun_out/r2w_pytest.log 2>&1; echo "pytest rc=$?"
grep -E "^(FAILED|ERROR)|^E  |passed|failed" gpurun_out/r2w_pytest.log | tail -5
python bench.py --workload c5 --steps 3 --warmup 3 --no-cpu-baseline --no-fast > gpurun_out/r2w_bench_c5_e2e.log 2>&1; echo "rc=$?"
tail -1 gpurun_out/r2w_bench_c5_e2e.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('value %.4g ms %.2f e2e %.4g e2e ms %.2f single %.1f frac %.4f'%(d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['e2e']['single_shot_ms'], r['frac']))"
