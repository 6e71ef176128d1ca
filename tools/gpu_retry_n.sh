#!/bin/bash
# usage: tools/gpu_retry_n.sh <gpus> <timeout_s> '<command>'   -- retries while gpurun has no slot (exit 3)
G=$1; shift; T=$1; shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --gpus $G --timeout $T -- "$@"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 120
done
exit 3
