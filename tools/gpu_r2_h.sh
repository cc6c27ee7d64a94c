#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "compaction" 2>&1 | tail -5
for pb in 0 16 24; do for tr in 1250000 10000000; do
  NI_B200_PARK_BELOW=$pb python bench.py --traj $tr --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-fast > gpurun_out/q.log 2>&1
  echo "park_below $pb traj $tr rc=$? $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/q.log | head -1) $(grep -o '"kernel_ms": [0-9.]*' gpurun_out/q.log | head -1) $(grep -o '"frac": [0-9.]*' gpurun_out/q.log | head -1)"
done; done
for pb in 0 16; do
  NI_B200_PARK_BELOW=$pb python bench.py --workload c4 --steps 3 --warmup 2 --no-cpu-baseline --no-e2e --no-fast > gpurun_out/q.log 2>&1
  echo "c4 park_below $pb rc=$? $(grep -o '"kernel_ms": [0-9.]*' gpurun_out/q.log | head -1) $(grep -o '"frac": [0-9.]*' gpurun_out/q.log | head -1)"
done
