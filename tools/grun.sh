#!/bin/bash
# usage: grun.sh <logfile> <timeout> <cmd...>   -- retries while the pod answers "transient/busy"
log=$1; shift; to=$1; shift
for i in $(seq 1 40); do
  gpurun --timeout $to -- "$@" > $log 2>&1
  rc=$?
  if grep -q "status=transient\|retry in a few minutes" $log || [ $rc -eq 3 ]; then sleep 45; continue; fi
  break
done
echo "grun done rc=$rc" >> $log
