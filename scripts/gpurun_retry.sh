#!/bin/bash
# usage: scripts/gpurun_retry.sh <gpus> <timeout> '<command>'   -- retries while the pod is busy (exit 3)
G=$1; T=$2; shift 2
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --gpus $G --timeout $T -- "$@"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 90
done
exit 3
