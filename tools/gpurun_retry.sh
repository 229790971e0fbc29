#!/bin/bash
# tools/gpurun_retry.sh [gpurun options] -- 'command' : retry while the pod answers "transient" (no slot free)
for i in $(seq 1 30); do
  out=$(/usr/local/graft/bin/gpurun "$@" 2>&1)
  echo "$out" | tail -40
  if echo "$out" | grep -q "status=transient"; then sleep 45; continue; fi
  break
done
