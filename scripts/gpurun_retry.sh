#!/bin/bash
# usage: scripts/gpurun_retry.sh <timeout_s> [--gpus N] -- '<command>'   -- retries while the pod answers "transient"/busy
T=$1; shift
for i in $(seq 1 12); do
  OUT=$(/usr/local/graft/bin/gpurun --timeout $T "$@" 2>&1)
  echo "$OUT" | tail -150
  if echo "$OUT" | grep -q "status=transient\|exit code 3\|rc=3"; then echo "[retry $i] pod busy, sleeping"; sleep 45; continue; fi
  break
done
