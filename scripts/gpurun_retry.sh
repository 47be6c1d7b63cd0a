#!/bin/bash
# scripts/gpurun_retry.sh <log> <gpurun args...>: retries while the pod answers "busy" (exit code 3)
LOG=$1; shift
for i in $(seq 1 30); do
  gpurun "$@" > "$LOG" 2>&1; rc=$?
  if [ $rc -ne 3 ] && ! grep -q "status=transient" "$LOG"; then exit $rc; fi
  sleep 90
done
exit 3
