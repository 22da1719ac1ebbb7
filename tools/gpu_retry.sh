#!/bin/bash
# development helper: retry a gpurun call while the pool answers "busy" (exit 3)
# usage: tools/gpu_retry.sh <gpurun args...>
for attempt in $(seq 1 20); do
  /usr/local/graft/bin/gpurun "$@"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 120
done
exit 3
