#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_full_size.py tests/test_gpu_api.py -q -k "heat or burgers or stencil" > gpurun_out/pytest_large.log 2>&1; echo "pytest rc=$?"
grep -E "^(FAILED|ERROR)|^E  |passed|failed" gpurun_out/pytest_large.log | tail -8
timeout 300 python bench.py --workload c5 --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.readlines()[-1]); print('c5', round(d['value']/1e6,2), round(d['roofline']['kernel_ms'],2), round(d['roofline']['frac'],4), 'e2e', round(d['e2e']['value']/1e6,2))"
