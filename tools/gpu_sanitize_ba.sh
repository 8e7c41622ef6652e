#!/bin/bash
# compute-sanitizer memcheck over the BA, Cholesky, search and matcher kernels (the extractor's logs are profiles/r2c_sanitizer_*.log):
# the reference-function fixtures and the small oracle cases; the BASELINE-size configs are left out (memcheck is ~30x slower).
mkdir -p gpurun_out
TAG=${TAG:-r2d}
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_ba_gpu.py tests/test_ba_leaf_gpu.py -m gpu -q \
  -k "reference_function or static_ba or dynamic_ba_matches or pose_optimization_matches or out_of_range or stop_flag or degenerate or leaf or huber" \
  > gpurun_out/${TAG}_sanitizer_memcheck_ba.log 2>&1; echo "ba rc=$?"; tail -3 gpurun_out/${TAG}_sanitizer_memcheck_ba.log
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_chol_gpu.py tests/test_search_gpu.py tests/test_match_gpu.py -m gpu -q \
  > gpurun_out/${TAG}_sanitizer_memcheck_search_match_chol.log 2>&1; echo "search rc=$?"; tail -3 gpurun_out/${TAG}_sanitizer_memcheck_search_match_chol.log
