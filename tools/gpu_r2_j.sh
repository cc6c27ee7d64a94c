#!/bin/bash
# round 2, call J: short division + lite guard + carried signed step -- device math check, parity, A/B against the base library
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_api.py tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/r2j_pytest.log 2>&1; echo "pytest rc=$?"
grep -E "^(FAILED|ERROR)|^E  |passed|failed" gpurun_out/r2j_pytest.log | tail -12
bash tools/exp_c3.sh --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-fast 2>&1 | grep "=="
bash tools/exp_c3.sh --workload c4 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-fast 2>&1 | grep "=="
bash tools/exp_c3.sh --workload c2 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-fast 2>&1 | grep "=="
