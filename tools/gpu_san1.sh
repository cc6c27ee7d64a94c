#!/bin/bash
mkdir -p gpurun_out
python tools/sanitize_case.py > gpurun_out/san_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/san_plain.log; exit 1; }
cat gpurun_out/san_plain.log
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 3 python tools/sanitize_case.py > gpurun_out/san_memcheck.log 2>&1; echo "memcheck rc=$?"
grep -E "ERROR SUMMARY|Invalid|out of bounds|misaligned" gpurun_out/san_memcheck.log | head -10
