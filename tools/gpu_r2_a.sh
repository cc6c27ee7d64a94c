#!/bin/bash
# round 2, call A: issue-slot microbenchmark, corrected FP64 peak (+ ncu pipe utilisation of the peak kernel), baseline bench
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r2a_gpu.txt
./gpurun_exp/bin/fp64_shadow > gpurun_out/r2a_fp64_shadow.txt 2>&1
echo "shadow rc=$?"
python - <<'PY' > gpurun_out/r2a_peak.txt 2>&1
import ctypes as C
from numba_integrators_b200 import _lib
pk = C.c_double(0)
for it in (2000, 4000):
    _lib.check(_lib.lib().ni_fp64_peak(0, it, 3, pk)); print('ni_fp64_peak iters', it, pk.value, 'TFLOP/s')
PY
cat gpurun_out/r2a_peak.txt
ncu --metrics sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed,gpu__time_duration.sum,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:ni_k_dfma_peak -c 2 --csv --log-file gpurun_out/r2a_peak_ncu.csv python -c "
import ctypes as C
from numba_integrators_b200 import _lib
pk = C.c_double(0); _lib.check(_lib.lib().ni_fp64_peak(0, 2000, 1, pk))" > gpurun_out/r2a_peak_ncu.log 2>&1
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2a_bench_c3.log 2>&1; tail -1 gpurun_out/r2a_bench_c3.log | cut -c1-400
