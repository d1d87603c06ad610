#!/bin/bash
# usage: tools/gpurun_retry.sh <timeout_s> [--gpus N] '<command>'   -- retries while the pod answers busy / transient (exit 3)
T=$1; shift
EXTRA=()
if [ "$1" = "--gpus" ]; then EXTRA=(--gpus "$2"); shift 2; fi
for i in $(seq 1 20); do
  /usr/local/graft/bin/gpurun --timeout "$T" "${EXTRA[@]}" -- "$@"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 60
done
exit 3
