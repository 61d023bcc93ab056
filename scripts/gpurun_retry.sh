#!/bin/bash
# usage: gpurun_retry.sh <logfile> <timeout_s> <command...>   (retries while the pod answers busy / transient)
LOG=$1; shift; TO=$1; shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout $TO -- "$@" > $LOG 2>&1
  rc=$?
  if grep -q "status=transient\|retry in a few minutes\|another call" $LOG || [ $rc -eq 3 ]; then sleep 90; continue; fi
  break
done
echo "gpurun_retry finished rc=$rc" >> $LOG
