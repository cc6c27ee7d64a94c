#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/exp3.txt
for cfg in "-DNI_LARGE_MAX_THREADS=512" "-DNI_LARGE_MAX_THREADS=1024" "-DNI_LARGE_MAX_THREADS=512 -DNI_LARGE_K2_IN_REG=0"; do
  NI_EXTRA_NVCC_FLAGS="$cfg" python -m numba_integrators_b200.build --force -v 2>&1 | grep -A2 "ni_k_large_runI13NiStencilHeatLi1ELb0ELb0ELi[48]E" | grep -E "Used|spill" | tr '\n' ' ' >> gpurun_out/exp3.txt
  r=$(python bench.py --workload c5 --steps 3 --warmup 2 --no-e2e --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.readlines()[-1]); print(round(d['value']/1e6,2), round(d['roofline']['kernel_ms'],2), round(d['roofline']['frac'],4))")
  echo "$cfg : Msteps/s kernel_ms frac = $r" | tee -a gpurun_out/exp3.txt
done
python -m pytest tests/test_gpu_parity.py -q -k "heat or burgers" 2>&1 | tail -2 | tee -a gpurun_out/exp3.txt
