#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_api.py -m gpu -q -x -k "terminal or gather" 2>&1 | tail -3
python bench.py --steps 5 --warmup 3 --cpu-seconds 4 > gpurun_out/r2g_bench.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/r2g_bench.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('value %.4g e2e %.4g ms %.2f frac %.4f fast %s launches %s'%(d['value'], d['e2e']['value'], d['ms_per_step'], r['frac'], r.get('fast'), d['gpu_launches']))"
