#!/bin/bash
# gpurun with retries on "busy" (exit 3 / transient): usage gpurun_retry.sh LOGFILE TIMEOUT 'command'
log=$1; to=$2; shift 2
for i in 1 2 3 4 5 6 7 8 9 10; do
  /usr/local/graft/bin/gpurun --timeout $to -- "$@" > $log 2>&1
  if grep -q "status=transient\|status=busy" $log; then sleep 90; continue; fi
  break
done
