#!/bin/bash
# usage: tools/gpurun_retry.sh [--gpus N] --timeout S -- 'command'   (retries while the pod answers "transient")
for attempt in 1 2 3 4 5 6 7 8 9 10 11 12; do
  out=$(/usr/local/graft/bin/gpurun "$@" 2>&1)
  if echo "$out" | grep -q "status=transient"; then
    sleep 150
    continue
  fi
  echo "$out"
  exit 0
done
echo "$out"
exit 3
