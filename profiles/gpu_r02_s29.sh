#!/bin/bash
# round 2, session 29: compute-sanitizer over the final build (CPSO bound decision, split VD-CMA units, concurrent seeds)
tag=r02s29
mkdir -p gpurun_out
( timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python profiles/sanitize_small.py 2>&1 | tail -60 ) > gpurun_out/${tag}_memcheck.log
grep -E "ERROR SUMMARY" gpurun_out/${tag}_memcheck.log
( timeout 900 compute-sanitizer --tool racecheck --print-limit 20 python profiles/sanitize_small.py 2>&1 | tail -60 ) > gpurun_out/${tag}_racecheck.log
grep -E "RACECHECK SUMMARY" gpurun_out/${tag}_racecheck.log
( timeout 600 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests/test_parallel.py -m gpu -q -k "concurrent" 2>&1 | tail -12 ) > gpurun_out/${tag}_memcheck_concurrent_seeds.log
grep -E "ERROR SUMMARY|passed|failed" gpurun_out/${tag}_memcheck_concurrent_seeds.log
