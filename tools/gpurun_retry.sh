#!/bin/bash
# local helper: run a gpurun call, retrying while the pod answers "transient" (nothing charged)
#   usage: tools/gpurun_retry.sh <timeout_s> <command...>
T=$1; shift
for i in $(seq 1 12); do
  out=$(/usr/local/graft/bin/gpurun --timeout $T -- "$@" 2>&1)
  if echo "$out" | grep -q "status=transient"; then sleep 45; continue; fi
  echo "$out" | tail -40
  exit 0
done
echo "gave up after 12 transient answers"
