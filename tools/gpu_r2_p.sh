#!/bin/bash
# round 2, call P: A/B of the FP64-compare absmax / clamp variant against the product library
mkdir -p gpurun_out
bash tools/exp_c3.sh --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-fast 2>&1 | grep "=="
bash tools/exp_c3.sh --workload c4 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-fast 2>&1 | grep "=="
bash tools/exp_c3.sh --workload c2 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-fast 2>&1 | grep "=="
