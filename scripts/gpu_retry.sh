#!/bin/bash
# usage: scripts/gpu_retry.sh <logfile> <timeout> <command...>   -- retries while the pod has no free GPU slot
log=$1; shift; to=$1; shift
for i in $(seq 1 40); do
  gpurun --timeout $to -- "$@" > $log.try 2>&1
  cat $log.try >> $log.all
  if grep -q "status=transient" $log.try; then sleep 45; continue; fi
  break
done
mv $log.try $log
