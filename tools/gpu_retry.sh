#!/bin/bash
# tools/gpu_retry.sh <logfile> <timeout> <command...>: gpurun with retries on "no slot" (exit 3) / transient answers.
log=$1; to=$2; shift 2
for attempt in 1 2 3 4 5 6 7 8; do
  /usr/local/graft/bin/gpurun --timeout $to -- "$@" > $log 2>&1
  rc=$?
  if grep -q "status=transient\|status=busy" $log || [ $rc -eq 3 ]; then sleep 90; continue; fi
  break
done
exit $rc
