#!/bin/bash
# gpu_retry.sh <log> <timeout> <command...>: gpurun with retries while the pod answers "no slot right now" (nothing is charged for those)
log=$1; to=$2; shift 2
for i in $(seq 1 40); do
  gpurun --timeout $to -- "$@" > $log 2>&1
  if grep -q "status=transient\|retry in a few minutes" $log; then sleep 45; continue; fi
  break
done
