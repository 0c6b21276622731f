#!/bin/bash
# usage: profiles/gpu_retry.sh <timeout-seconds> '<command>'   -- gpurun with retries while the pod has no free GPU slot (exit code 3)
t=$1; shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout "$t" -- "$@"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 45
done
exit 3
