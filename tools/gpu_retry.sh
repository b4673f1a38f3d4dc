#!/bin/bash
# usage: tools/gpu_retry.sh <timeout> '<command>'  -- retries while the pod reports "transient" (busy), up to ~40 min
for i in $(seq 1 14); do
  out=$(/usr/local/graft/bin/gpurun --timeout "$1" -- "$2" 2>&1)
  if echo "$out" | grep -q "status=transient"; then sleep 150; continue; fi
  echo "$out"; exit 0
done
echo "$out"; exit 3
