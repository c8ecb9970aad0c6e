#!/bin/bash
# usage: tools/gpurun_retry.sh <timeout_s> <logfile> [--gpus N] -- <command>   (retries while the pod answers busy / transient)
T=$1; LOG=$2; shift 2
for i in $(seq 1 30); do
  /usr/local/graft/bin/gpurun --timeout $T "$@" > $LOG 2>&1
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 120
done
exit 3
