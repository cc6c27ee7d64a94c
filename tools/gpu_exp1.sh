#!/bin/bash
# occupancy / block-size experiment on the C3 run kernel (rebuilds on the box)
mkdir -p gpurun_out
: > gpurun_out/exp1.txt
for cfg in "-DNI_MIN_BLOCKS=1" "-DNI_MIN_BLOCKS=6" "-DNI_MIN_BLOCKS=7" "-DNI_MIN_BLOCKS=8" "-DNI_BLOCK=256 -DNI_MIN_BLOCKS=3" "-DNI_BLOCK=64 -DNI_MIN_BLOCKS=12"; do
  rm -f numba_integrators_b200/csrc/build/ni_builtin_lorenz63.o
  NI_EXTRA_NVCC_FLAGS="$cfg" python -m numba_integrators_b200.build --force > /dev/null 2>&1 || { echo "build failed $cfg" >> gpurun_out/exp1.txt; continue; }
  r=$(python bench.py --traj 4000000 --steps 3 --warmup 2 --no-e2e --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.readlines()[-1]); print(d['value']/1e9, d['roofline']['kernel_ms'], d['roofline']['frac'])")
  echo "$cfg : Gsteps/s kernel_ms frac = $r" | tee -a gpurun_out/exp1.txt
done
