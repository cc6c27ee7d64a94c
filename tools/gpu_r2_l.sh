#!/bin/bash
# round 2, call L: record-slot look-ahead of the CTA-per-trajectory kernels -- tests, C5 with recording, C3 / C4 re-check
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_large_features.py tests/test_gpu_api.py -m gpu -q -x > gpurun_out/r2l_pytest.log 2>&1; echo "pytest rc=$?"
grep -E "^(FAILED|ERROR)|^E  |passed|failed" gpurun_out/r2l_pytest.log | tail -12
python bench.py --workload c5 --record --traj 8192 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/r2l_bench_c5_rec.log 2>&1; echo "c5 rec rc=$? $(grep -o '"value": [0-9.e+]*' gpurun_out/r2l_bench_c5_rec.log | head -1) $(grep -o '"hbm": {[^}]*}' gpurun_out/r2l_bench_c5_rec.log | cut -c1-120)"
for wk in c3 c4; do python bench.py --workload $wk --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2l_bench_$wk.log 2>&1; echo "$wk rc=$? $(grep -o '"kernel_ms": [0-9.e+]*' gpurun_out/r2l_bench_$wk.log | head -1) $(grep -o '"frac": [0-9.]*' gpurun_out/r2l_bench_$wk.log | head -3 | tr '\n' ' ')"; done
