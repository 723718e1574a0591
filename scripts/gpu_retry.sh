#!/bin/bash
# usage: gpu_retry.sh <log> <timeout> [--gpus N] <command>  -- retries while the pod answers busy (nothing charged)
log=$1; shift; to=$1; shift
opts=""
if [ "$1" == "--gpus" ]; then opts="--gpus $2"; shift; shift; fi
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout $to $opts -- "$@" > $log 2>&1
  if grep -q "nothing was charged" $log; then sleep 90; continue; fi
  break
done
