#!/bin/bash
# usage: tools/gpurun_retry.sh <out-file> <gpurun args...>   - retries while the pod answers "busy" (exit 3), up to ~40 min
out=$1; shift
for i in $(seq 1 14); do
  /usr/local/graft/bin/gpurun "$@" > "$out" 2>&1
  rc=$?
  if [ $rc -ne 3 ] && ! grep -q "status=transient" "$out"; then exit $rc; fi
  sleep 150
done
exit 3
