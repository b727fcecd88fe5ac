#!/bin/bash
# usage: tools/gpu.sh <timeout-seconds> <log-file> <command...>   -- gpurun with retries while no box is free (exit code 3)
T=$1; LOG=$2; shift 2
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout "$T" -- "$@" > "$LOG" 2>&1
  rc=$?
  if [ $rc -ne 3 ]; then echo "gpurun rc=$rc (attempt $i)" >> "$LOG"; exit $rc; fi
  sleep 45
done
echo "gpurun: no box after 40 attempts" >> "$LOG"; exit 3
