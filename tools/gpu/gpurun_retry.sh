#!/bin/bash
# usage: [GPUS=2] tools/gpu/gpurun_retry.sh <timeout> <command...>   -- retries while the pod's GPU slots are busy
t=$1; shift
for i in 1 2 3 4 5 6 7 8; do
  out=$(/usr/local/graft/bin/gpurun ${GPUS:+--gpus $GPUS} --timeout $t -- "$@" 2>&1)
  echo "$out"
  if ! echo "$out" | grep -q "status=transient"; then break; fi
  sleep 60
done
