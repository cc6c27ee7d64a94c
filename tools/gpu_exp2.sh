#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/exp2.txt
for k in 4 5 6; do
  NI_EXTRA_NVCC_FLAGS="-DNI_MIN_BLOCKS(n)=$k" python -m numba_integrators_b200.build --force > /dev/null 2>&1 || { echo "build failed $k" >> gpurun_out/exp2.txt; continue; }
  for ar in fast strict; do
  r=$(python bench.py --arithmetic $ar --steps 3 --warmup 2 --no-e2e --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.readlines()[-1]); print(round(d['value']/1e9,2), round(d['roofline']['kernel_ms'],2), round(d['roofline']['frac'],4))")
  echo "min_blocks=$k $ar : Gsteps/s kernel_ms frac = $r" | tee -a gpurun_out/exp2.txt
  done
done
