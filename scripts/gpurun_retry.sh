#!/bin/bash
# retry a gpurun call while the pod answers "no slot" (exit code 3); usage: gpurun_retry.sh <tries> <gpurun args...>
tries=$1; shift
for i in $(seq 1 $tries); do
  /usr/local/graft/bin/gpurun "$@"; rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 150
done
exit 3
