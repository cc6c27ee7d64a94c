#!/bin/bash
# round 2, call V: full GPU suite on the final tree + default bench
mkdir -p gpurun_out
timeout 1700 python -m pytest tests -m gpu -q > gpurun_out/r2v_pytest_gpu.log 2>&1; echo "pytest rc=$?"
grep -E "^(FAILED|ERROR)|^E  |passed|failed" gpurun_out/r2v_pytest_gpu.log | tail -12
python __graft_entry__.py smoke 2>&1 | tail -1
( time python bench.py > gpurun_out/r2v_bench_c3.log 2>&1 ) 2>&1 | grep real; echo "c3 rc=$?"; tail -1 gpurun_out/r2v_bench_c3.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('value %.4g e2e %.4g (single shot %.1f ms) ms %.2f frac %.4f fast %.4f ordered %.4g launches %s cpu %.4g'%(d['value'], d['e2e']['value'], d['e2e']['single_shot_ms'], d['ms_per_step'], r['frac'], r['fast']['frac'], d['cost_ordered']['value'], d['gpu_launches'], d['cpu_baseline']['value']))"
