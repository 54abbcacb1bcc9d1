#!/bin/bash
# gpurun with retries on "busy / transient" answers (nothing is charged for those).  usage: gpurun_retry.sh <log> <timeout> [--gpus N] -- <cmd>
log=$1; shift; to=$1; shift
for attempt in 1 2 3 4 5 6 7 8 9 10 11 12; do
  /usr/local/graft/bin/gpurun --timeout $to "$@" > $log 2>&1
  if grep -q "status=transient\|rc=3\|no box\|busy" $log && ! grep -q "status=ok" $log; then
    echo "attempt $attempt: transient, retrying in 120 s" >> ${log}.retries; sleep 120
  else
    break
  fi
done
