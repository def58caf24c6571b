#!/bin/bash
# usage: tools/gpurun_retry.sh <gpus> <timeout> <command...>   -- retries while the pod has no free slot (exit 3)
G=$1; T=$2; shift 2
for k in 1 2 3 4 5 6 7 8; do
  if [ "$G" = "1" ]; then /usr/local/graft/bin/gpurun --timeout $T -- "$@"; else /usr/local/graft/bin/gpurun --gpus $G --timeout $T -- "$@"; fi
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 150
done
exit 3
