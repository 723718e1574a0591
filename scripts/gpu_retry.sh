#!/bin/bash
# usage: gpu_retry.sh <log> <timeout> <command...>  -- retries while the pod answers busy (exit 3 / transient with nothing charged)
log=$1; shift; to=$1; shift
for i in $(seq 1 30); do
  /usr/local/graft/bin/gpurun --timeout $to -- "$@" > $log 2>&1
  if grep -q "nothing was charged" $log; then sleep 90; continue; fi
  break
done
