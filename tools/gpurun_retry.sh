#!/bin/bash
# usage: tools/gpurun_retry.sh <logfile> <gpurun args...>   -- retries while the pod answers "busy" (exit code 3 / transient)
log=$1; shift
for i in $(seq 1 20); do
  /usr/local/graft/bin/gpurun "$@" > "$log" 2>&1
  if ! grep -q "status=transient\|no box or slot" "$log"; then exit 0; fi
  sleep 120
done
