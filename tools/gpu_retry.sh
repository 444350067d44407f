#!/bin/bash
# Retry a gpurun call while the pod answers "transient" (exit code 3): usage  tools/gpu_retry.sh <log> <timeout_s> <command...>
LOG=$1; shift; T=$1; shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout $T -- "$@" > $LOG 2>&1
  rc=$?
  if [ $rc -ne 3 ] && ! grep -q "status=transient" $LOG; then exit $rc; fi
  sleep 90
done
