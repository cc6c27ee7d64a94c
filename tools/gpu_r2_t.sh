#!/bin/bash
# round 2, call T: launch order (longest first) -- test, and its effect on C3 (1.25 M and 10 M trajectories) and C4
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_api.py -m gpu -q -x -k "launch_order" > gpurun_out/r2t_pytest.log 2>&1; echo "pytest rc=$?"
grep -E "^(FAILED|ERROR)|^E  |passed|failed" gpurun_out/r2t_pytest.log | tail -8
for args in "--traj 1250000" "" "--workload c4"; do
  ( time python bench.py $args --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-fast > gpurun_out/q.log 2>&1 ) 2>&1 | grep real | tr '\n' ' '
  echo "[$args] rc=$? $(grep -o '"kernel_ms": [0-9.]*' gpurun_out/q.log | head -1) ordered: $(grep -o '"cost_ordered": {[^}]*' gpurun_out/q.log | cut -c1-120)"
done
