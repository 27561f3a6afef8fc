#!/bin/bash
# usage: tools/gpurun_retry.sh <timeout-seconds> <command...>   -- retries while the pod has no free GPU slot (exit code 3)
t=$1; shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout "$t" -- "$@" > /tmp/gpurun_last.log 2>&1
  rc=$?
  if grep -q "status=transient" /tmp/gpurun_last.log || [ $rc -eq 3 ]; then sleep 90; continue; fi
  break
done
tail -60 /tmp/gpurun_last.log
exit $rc
