#!/bin/bash
# gpu_retry_n.sh <gpus> <log> <timeout> <command...>: like gpu_retry.sh for an N-GPU box
n=$1; log=$2; to=$3; shift 3
for i in $(seq 1 30); do
  gpurun --gpus $n --timeout $to -- "$@" > $log 2>&1
  if grep -q "status=transient\|retry in a few minutes\|status=busy" $log; then sleep 60; continue; fi
  break
done
