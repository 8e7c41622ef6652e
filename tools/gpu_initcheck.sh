#!/bin/bash
# compute-sanitizer initcheck (reads of uninitialised device memory) and synccheck (divergent barriers) over small cases of
# every kernel family.  The full report is large (one record per thread), so only the per-kernel / per-line aggregate is kept.
mkdir -p gpurun_out
TAG=${TAG:-r2h}
SEL=${SEL:-'reference_function or reference_operator or getters or edge_cases or (static_ba and (tiny or small)) or pose_optimization_matches or dense_solve_matches or best2 or stereo_no_matches or distinctive'}
TOOLS=${TOOLS:-'initcheck synccheck'}
for tool in $TOOLS; do
  log=gpurun_out/${TAG}_sanitizer_${tool}
  timeout ${LIMIT:-420} compute-sanitizer --tool $tool --error-exitcode 9 --show-backtrace device --print-limit 200000 \
    python -m pytest tests -m gpu -q -k "$SEL" > $log.full 2>&1; echo "$tool rc=$?"
  { grep -E "passed|failed|ERROR SUMMARY" $log.full
    echo "--- records by kernel and source line"
    grep -E "Device Frame| at .* in " $log.full | sed -E 's/0x[0-9a-f]+//g; s/\+ in / in /; s/\(.*\) *(const)?//' | sort | uniq -c | sort -rn | head -40
    echo "--- head of the full report"
    grep -v "^\.*$" $log.full | head -60
  } > $log.log
  rm -f $log.full
  cat $log.log
done
