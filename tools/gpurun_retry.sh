#!/bin/bash
# gpurun with retries while the pod has no slot (status transient / exit code 3): tools/gpurun_retry.sh <timeout> '<command>'
t=$1; shift
for i in 1 2 3 4 5 6 7 8; do
  /usr/local/graft/bin/gpurun --timeout $t -- "$@" > /tmp/gpurun_last.log 2>&1
  rc=$?
  if grep -q "status=transient" /tmp/gpurun_last.log || [ $rc -eq 3 ]; then sleep 120; continue; fi
  break
done
cat /tmp/gpurun_last.log
