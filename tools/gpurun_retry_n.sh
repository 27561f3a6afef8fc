#!/bin/bash
# usage: tools/gpurun_retry_n.sh <gpus> <timeout-seconds> <command...>   -- multi-GPU form of tools/gpurun_retry.sh
n=$1; t=$2; shift; shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --gpus "$n" --timeout "$t" -- "$@" > /tmp/gpurun_last.log 2>&1
  rc=$?
  if grep -q "status=transient" /tmp/gpurun_last.log || [ $rc -eq 3 ]; then sleep 90; continue; fi
  break
done
tail -60 /tmp/gpurun_last.log
exit $rc
