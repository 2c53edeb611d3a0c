#!/bin/bash
# gpurun with retries while the pod answers "transient" (nothing is charged for those).  Usage: gpurun_retry.sh <out-file> <gpurun args...>
OUT=$1; shift
for i in $(seq 1 20); do
  /usr/local/graft/bin/gpurun "$@" > "$OUT" 2>&1
  if ! grep -q "status=transient" "$OUT"; then exit 0; fi
  sleep 90
done
exit 3
