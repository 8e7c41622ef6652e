#!/bin/bash
# retries a gpurun call while the pod answers "transient" (nothing charged); usage: tools/gpurun_retry.sh <log> <timeout> <command>
LOG=$1; TO=$2; shift 2
for i in $(seq 1 12); do
  /usr/local/graft/bin/gpurun --timeout $TO -- "$@" > $LOG 2>&1
  grep -q "status=transient" $LOG || exit 0
  sleep 120
done
