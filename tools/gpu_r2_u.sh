#!/bin/bash
# round 2, call U: A/B of the lazy (batched) refill variants
mkdir -p gpurun_out
bash tools/exp_c3.sh --steps 4 --warmup 3 --no-cpu-baseline --no-e2e --no-fast 2>&1 | grep "=="
bash tools/exp_c3.sh --traj 1250000 --steps 4 --warmup 3 --no-cpu-baseline --no-e2e --no-fast 2>&1 | grep "=="
bash tools/exp_c3.sh --workload c4 --steps 3 --warmup 2 --no-cpu-baseline --no-e2e --no-fast 2>&1 | grep "=="
