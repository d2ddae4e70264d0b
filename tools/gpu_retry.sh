#!/bin/bash
# usage: tools/gpu_retry.sh <gpus> <timeout_s> '<command>'   — retries while the pod answers busy (exit 3 / transient)
G=$1; T=$2; CMD=$3
for i in $(seq 1 30); do
  if [ "$G" = "1" ]; then
    /usr/local/graft/bin/gpurun --timeout $T -- "$CMD" > /tmp/gpurun_last.log 2>&1
  else
    /usr/local/graft/bin/gpurun --gpus $G --timeout $T -- "$CMD" > /tmp/gpurun_last.log 2>&1
  fi
  rc=$?
  if grep -q "status=transient\|nothing was charged" /tmp/gpurun_last.log || [ $rc -eq 3 ]; then
    sleep 60; continue
  fi
  break
done
tail -40 /tmp/gpurun_last.log
exit $rc
