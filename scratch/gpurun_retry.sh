#!/bin/bash
# gpurun with retries while the pod answers "transient" (nothing charged): scratch/gpurun_retry.sh <timeout> <command...>
T=$1; shift
for i in $(seq 1 12); do
  OUT=$(/usr/local/graft/bin/gpurun --timeout $T -- "$@" 2>&1)
  echo "$OUT" | tail -4
  if ! echo "$OUT" | grep -q "status=transient"; then exit 0; fi
  sleep 100
done
