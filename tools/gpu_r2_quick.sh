#!/bin/bash
# quick check of a kernel change: parity tests, then the product library against gpurun_exp/libni_*.so on C3 (strict),
# then C3 fast / C4 / C2 of the product library
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py tests/test_gpu_full_size.py tests/test_gpu_fuzz.py -m gpu -q -x 2>&1 | tail -3
bash tools/exp_c3.sh --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-fast 2>&1 | grep "=="
for a in "--arithmetic fast" "--workload c4" "--workload c2" "--workload c4 --arithmetic fast" "--workload c5"; do
  python bench.py $a --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-fast > gpurun_out/q.log 2>&1
  echo "$a rc=$? $(grep -o '"value": [0-9.e+]*' gpurun_out/q.log | head -1) $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/q.log | head -1) $(grep -o '"frac": [0-9.]*' gpurun_out/q.log | head -1)"
done
