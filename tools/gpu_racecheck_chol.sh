#!/bin/bash
mkdir -p gpurun_out
TAG=${TAG:-r2f}
timeout 1200 compute-sanitizer --tool racecheck --print-limit 400 --error-exitcode 9 python -m pytest tests/test_ba_gpu.py tests/test_chol_gpu.py -m gpu -q \
  -k "reference_function or (static_ba and (tiny or small)) or pose_optimization_matches or dense_solve_matches" \
  > gpurun_out/${TAG}_sanitizer_racecheck_ba.log 2>&1; echo "ba rc=$?"; tail -4 gpurun_out/${TAG}_sanitizer_racecheck_ba.log
grep -E "Error:|Warning:" gpurun_out/${TAG}_sanitizer_racecheck_ba.log | sed 's/+0x[0-9a-f]*//g' | sort | uniq -c | sort -rn | head -20
