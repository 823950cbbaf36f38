#!/bin/bash
# usage: [GPUS=N] grun.sh <logfile> <timeout> <cmd...>   -- retries while the pod answers "transient/busy"
log=$1; shift; to=$1; shift
gp=""; [ -n "$GPUS" ] && gp="--gpus $GPUS"
for i in $(seq 1 60); do
  gpurun $gp --timeout $to -- "$@" > $log 2>&1
  rc=$?
  if grep -q "status=transient\|retry in a few minutes\|another call" $log || [ $rc -eq 3 ] || [ $rc -eq 2 ]; then sleep 45; continue; fi
  break
done
echo "grun done rc=$rc" >> $log
