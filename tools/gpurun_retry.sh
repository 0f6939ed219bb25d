#!/bin/bash
# usage: tools/gpurun_retry.sh <timeout> [--gpus N] -- '<command>'   (retries while the pod is busy: exit code 3)
for i in $(seq 1 30); do
  /usr/local/graft/bin/gpurun --timeout "$@"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 150
done
exit 3
