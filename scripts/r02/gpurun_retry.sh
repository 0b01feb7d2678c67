#!/bin/bash
# usage: gpurun_retry.sh <timeout_s> <logfile> [--gpus N] -- <command>
# retries while gpurun answers "busy" (exit 3), sleeping 2 min between tries (at most 12 tries)
T=$1; LOG=$2; shift 2
for i in $(seq 1 12); do
  /usr/local/graft/bin/gpurun --timeout $T "$@" > $LOG 2>&1; rc=$?
  [ $rc -ne 3 ] && exit $rc
  sleep 120
done
exit 3
