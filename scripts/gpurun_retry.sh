#!/bin/bash
# usage: scripts/gpurun_retry.sh <timeout-seconds> <command...>: retries while gpurun answers "busy" (exit 3)
T=$1; shift
for attempt in $(seq 1 20); do
  /usr/local/graft/bin/gpurun --timeout $T -- "$@"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  echo "[retry] attempt $attempt answered busy; sleeping 120 s"
  sleep 120
done
exit 3
