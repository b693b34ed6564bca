#!/bin/bash
# usage: scripts/gpu_retry.sh <log> <gpurun args...>   — retries while gpurun answers "busy" (exit 3)
LOG=$1; shift
for i in $(seq 1 30); do
  gpurun "$@" > "$LOG" 2>&1; rc=$?
  if [ $rc -ne 3 ]; then echo "exit $rc" >> "$LOG"; exit $rc; fi
  sleep 90
done
echo "exit 3 (gave up)" >> "$LOG"
