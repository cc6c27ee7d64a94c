#!/bin/bash
# round 2, call B: lean run body -- GPU test suite, then bench A/B against gpurun_exp/libni_*.so variants
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r2b_pytest_gpu.log 2>&1; echo "pytest rc=$?"
grep -E "^(FAILED|ERROR)|^E  |passed|failed" gpurun_out/r2b_pytest_gpu.log | tail -12
python __graft_entry__.py smoke 2>&1 | tail -1
bash tools/exp_c3.sh --steps 5 --warmup 3 --no-cpu-baseline --no-e2e 2>&1 | grep "=="
for wk in c4 c2 c5; do python bench.py --workload $wk --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2b_bench_$wk.log 2>&1; echo "$wk rc=$? $(grep -o '"value": [0-9.e+]*' gpurun_out/r2b_bench_$wk.log | head -1) $(grep -o '"frac": [0-9.]*' gpurun_out/r2b_bench_$wk.log | head -1)"; done
python bench.py --arithmetic fast --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2b_bench_c3_fast.log 2>&1; echo "c3 fast rc=$? $(grep -o '"value": [0-9.e+]*' gpurun_out/r2b_bench_c3_fast.log | head -1) $(grep -o '"frac": [0-9.]*' gpurun_out/r2b_bench_c3_fast.log | head -1)"
