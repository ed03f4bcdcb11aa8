#!/bin/bash
# usage: tools/gpu_retry.sh <timeout_s> '<command>'   -- retries while the pod answers "busy" (exit 3), up to 12 times
T=$1; shift
for i in $(seq 1 12); do
  /usr/local/graft/bin/gpurun --timeout "$T" -- "$@"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  echo "[gpu_retry] busy, attempt $i; sleeping 90 s"
  sleep 90
done
exit 3
