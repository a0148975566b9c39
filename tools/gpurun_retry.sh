#!/bin/bash
# usage: tools/gpurun_retry.sh <logfile> <timeout_s> <command...>   -- retries while the pod answers "transient"/busy
log=$1; shift; to=$1; shift
for attempt in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout "$to" -- "$@" > "$log" 2>&1
  rc=$?
  if grep -q "status=transient\|retry in a few minutes\|no box or slot" "$log" || [ $rc -eq 3 ]; then sleep 90; continue; fi
  break
done
echo "attempts=$attempt rc=$rc" >> "$log"
