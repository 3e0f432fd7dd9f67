#!/bin/bash
# gpurun with retries while the pod has no free GPU slot (exit code 3 = nothing charged)
#   scripts/gpurun_retry.sh [--gpus N] --timeout S -- <command>
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun "$@"; rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 45
done
exit 3
