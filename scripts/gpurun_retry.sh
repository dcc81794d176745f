#!/bin/bash
# usage: gpurun_retry.sh <timeout> <cmd...>   retries while the pod answers "transient"/busy (rc 3)
T="$1"; shift
for i in $(seq 1 40); do
  out=$(/usr/local/graft/bin/gpurun --timeout "$T" -- "$@" 2>&1)
  echo "$out" | tail -120
  if echo "$out" | grep -q "status=transient\|nothing was charged"; then echo "[retry $i] sleeping 90s"; sleep 90; continue; fi
  break
done
