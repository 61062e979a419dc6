#!/bin/bash
# gpurun with retries while the pod has no free slot (exit code 3: nothing charged):  scripts/gpurun_retry.sh [gpurun options] -- 'command'
for i in $(seq 12); do
  /usr/local/graft/bin/gpurun "$@"
  rc=$?
  [ $rc -ne 3 ] && exit $rc
  sleep 150
done
exit 3
