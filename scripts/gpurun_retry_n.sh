#!/bin/bash
# usage: gpurun_retry_n.sh <gpus> <logfile> <timeout_s> <command...>
G=$1; shift; LOG=$1; shift; TO=$1; shift
for i in $(seq 1 60); do
  /usr/local/graft/bin/gpurun --gpus $G --timeout $TO -- "$@" > $LOG 2>&1
  rc=$?
  if grep -q "status=transient\|retry in a few minutes\|another call\|busy" $LOG || [ $rc -eq 3 ]; then sleep 120; continue; fi
  break
done
echo "gpurun_retry finished rc=$rc" >> $LOG
