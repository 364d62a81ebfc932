#!/bin/bash
# usage: tools/gpurun_retry.sh <timeout-seconds> '<command>'   -- retries while the pod has no free slot (nothing is charged then)
T=$1; shift
for i in $(seq 1 40); do
  out=$(/usr/local/graft/bin/gpurun --timeout "$T" -- "$@" 2>&1)
  echo "$out" | tail -25
  if echo "$out" | grep -q "status=transient\|no box\|busy\|retry in a few"; then sleep 90; continue; fi
  break
done
