#!/bin/bash
# usage: tools/gpu_retry.sh <log> <timeout> [--gpus N] <command string> : retries while the pod answers busy (exit 3)
log=$1; to=$2; shift 2
extra=""
if [ "$1" = "--gpus" ]; then extra="--gpus $2"; shift 2; fi
for i in $(seq 1 30); do
  /usr/local/graft/bin/gpurun --timeout $to $extra -- "$@" > $log 2>&1
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 60
done
exit 3
