#!/bin/bash
# round 2, call K: full GPU suite on the new lean body, then C5 with recording (8192 systems: 43 GB log) and the other configs
mkdir -p gpurun_out
timeout 1700 python -m pytest tests -m gpu -q -x > gpurun_out/r2k_pytest_gpu.log 2>&1; echo "pytest rc=$?"
grep -E "^(FAILED|ERROR)|^E  |passed|failed" gpurun_out/r2k_pytest_gpu.log | tail -12
python __graft_entry__.py smoke 2>&1 | tail -1
for wk in c3 c4 c2 c5; do python bench.py --workload $wk --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2k_bench_$wk.log 2>&1; echo "$wk rc=$? $(grep -o '"value": [0-9.e+]*' gpurun_out/r2k_bench_$wk.log | head -1) $(grep -o '"frac": [0-9.]*' gpurun_out/r2k_bench_$wk.log | head -3 | tr '\n' ' ')"; done
python bench.py --workload c5 --record --traj 8192 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/r2k_bench_c5_rec.log 2>&1; echo "c5 rec rc=$? $(grep -o '"value": [0-9.e+]*' gpurun_out/r2k_bench_c5_rec.log | head -1) $(grep -o '"hbm": {[^}]*}' gpurun_out/r2k_bench_c5_rec.log | cut -c1-200)"
python bench.py --workload c5 --traj 8192 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/r2k_bench_c5_norec.log 2>&1; echo "c5 8192 norec rc=$? $(grep -o '"value": [0-9.e+]*' gpurun_out/r2k_bench_c5_norec.log | head -1)"
