#!/bin/bash
# round 2, call Q: full GPU suite on the final kernels, smoke, default bench (with e2e + cpu baseline) and the other configs
mkdir -p gpurun_out
timeout 1700 python -m pytest tests -m gpu -q > gpurun_out/r2q_pytest_gpu.log 2>&1; echo "pytest rc=$?"
grep -E "^(FAILED|ERROR)|^E  |passed|failed" gpurun_out/r2q_pytest_gpu.log | tail -12
python __graft_entry__.py smoke 2>&1 | tail -1
python bench.py > gpurun_out/r2q_bench_c3.log 2>&1; echo "c3 rc=$?"; tail -1 gpurun_out/r2q_bench_c3.log | cut -c1-400
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2q_bench_ref.log 2>&1; echo "ref rc=$?"; tail -1 gpurun_out/r2q_bench_ref.log | cut -c1-300
for wk in c4 c2 c5; do python bench.py --workload $wk --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2q_bench_$wk.log 2>&1; echo "$wk rc=$? $(grep -o '"value": [0-9.e+]*' gpurun_out/r2q_bench_$wk.log | head -1) $(grep -o '"frac": [0-9.]*' gpurun_out/r2q_bench_$wk.log | head -3 | tr '\n' ' ')"; done
python bench.py --workload c5 --record --traj 8192 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/r2q_bench_c5_rec.log 2>&1; echo "c5 rec rc=$? $(grep -o '"value": [0-9.e+]*' gpurun_out/r2q_bench_c5_rec.log | head -1)"
