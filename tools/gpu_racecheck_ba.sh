#!/bin/bash
# compute-sanitizer racecheck (shared-memory hazards) over the BA / pose / Cholesky / search / stereo kernels on small cases
mkdir -p gpurun_out
TAG=${TAG:-r2d}
timeout 1200 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_ba_gpu.py tests/test_chol_gpu.py -m gpu -q \
  -k "reference_function or (static_ba and (tiny or small)) or pose_optimization_matches or dense_solve_matches" \
  > gpurun_out/${TAG}_sanitizer_racecheck_ba.log 2>&1; echo "ba rc=$?"; tail -4 gpurun_out/${TAG}_sanitizer_racecheck_ba.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_search_gpu.py tests/test_match_gpu.py -m gpu -q -k "reference_function" \
  > gpurun_out/${TAG}_sanitizer_racecheck_search_match.log 2>&1; echo "search rc=$?"; tail -4 gpurun_out/${TAG}_sanitizer_racecheck_search_match.log
