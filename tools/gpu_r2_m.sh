#!/bin/bash
# round 2, call M: record rows of the CTA-per-trajectory kernels (chunked slots, padded transposition buffer)
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_large_features.py tests/test_gpu_api.py tests/test_gpu_scipy_semantics.py tests/test_gpu_fuzz.py -m gpu -q -x > gpurun_out/r2m_pytest.log 2>&1; echo "pytest rc=$?"
grep -E "^(FAILED|ERROR)|^E  |passed|failed" gpurun_out/r2m_pytest.log | tail -12
python bench.py --workload c5 --record --traj 8192 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/r2m_bench_c5_rec.log 2>&1; echo "c5 rec rc=$? $(grep -o '"value": [0-9.e+]*' gpurun_out/r2m_bench_c5_rec.log | head -1) $(grep -o '"hbm": {[^}]*}' gpurun_out/r2m_bench_c5_rec.log | cut -c1-120)"
python bench.py --workload c5 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/r2m_bench_c5.log 2>&1; echo "c5 rc=$? $(grep -o '"value": [0-9.e+]*' gpurun_out/r2m_bench_c5.log | head -1)"
