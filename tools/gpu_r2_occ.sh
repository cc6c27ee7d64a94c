#!/bin/bash
# throughput of the C3 run kernel against resident warps per scheduler (1 CTA of 128 threads per SM = 1 warp per scheduler)
mkdir -p gpurun_out
for b in 1 2 3 4 5; do
  NI_B200_BLOCKS_PER_SM=$b python bench.py --traj 1000000 --steps 3 --warmup 2 --no-cpu-baseline --no-e2e > gpurun_out/q.log 2>&1
  echo "blocks/SM $b rc=$? $(grep -o '"kernel_ms": [0-9.]*' gpurun_out/q.log | head -1) $(grep -o '"frac": [0-9.]*' gpurun_out/q.log | head -1)"
done
cp numba_integrators_b200/libni_b200.so /tmp/prod.so
cp gpurun_exp/libni_mb6.so numba_integrators_b200/libni_b200.so
python bench.py --traj 1000000 --steps 3 --warmup 2 --no-cpu-baseline --no-e2e > gpurun_out/q.log 2>&1
echo "blocks/SM 6 (80 regs) rc=$? $(grep -o '"kernel_ms": [0-9.]*' gpurun_out/q.log | head -1) $(grep -o '"frac": [0-9.]*' gpurun_out/q.log | head -1)"
cp /tmp/prod.so numba_integrators_b200/libni_b200.so
