#!/usr/bin/env bash
# usage: [GPUS=N] tools/gpurun_retry.sh <timeout> <logfile> <command...>   -- retries while gpurun answers "busy" (exit 3)
T=$1; LOG=$2; shift 2
G=()
if [ -n "${GPUS:-}" ]; then G=(--gpus "$GPUS"); fi
for i in $(seq 1 60); do
  /usr/local/graft/bin/gpurun --timeout "$T" "${G[@]}" -- "$@" > "$LOG" 2>&1
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 75
done
exit 3
