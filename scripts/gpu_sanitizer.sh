#!/bin/bash
# compute-sanitizer over the new code paths (small scenes): memcheck on the mapper + parity tests, racecheck on one mapper iteration
set -u
O=gpurun_out/r2san
mkdir -p $O
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 7 --print-limit 20 python -m pytest tests/test_gpu_mapper.py -m gpu -q -x -k "iteration_matches_packed_model or extend_appends or two_views_per_iteration" --timeout 1400 > $O/memcheck_mapper.log 2>&1
echo "memcheck mapper rc=$?" | tee -a $O/memcheck_mapper.log
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 --print-limit 20 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "test_forward_matches_oracle or empty_and_tiny or capacity_overflow" --timeout 800 > $O/memcheck_parity.log 2>&1
echo "memcheck parity rc=$?" | tee -a $O/memcheck_parity.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 7 --print-limit 20 python -m pytest tests/test_gpu_mapper.py -m gpu -q -x -k "iteration_matches_packed_model and 3" --timeout 800 > $O/racecheck_mapper.log 2>&1
echo "racecheck mapper rc=$?" | tee -a $O/racecheck_mapper.log
grep -h "ERROR SUMMARY\|RACECHECK SUMMARY\|passed\|failed" $O/*.log
