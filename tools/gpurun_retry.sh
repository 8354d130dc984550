#!/bin/bash
# usage: gpurun_retry.sh <logfile> <gpurun args...>  -- retries while the pod answers "no slot free" (nothing is charged for those)
log=$1; shift
for attempt in $(seq 1 20); do
  gpurun "$@" > "$log" 2>&1
  if ! grep -q "status=transient" "$log"; then exit 0; fi
  sleep 90
done
