#!/bin/bash
# usage: tools/gpu_retry.sh <timeout_s> '<command>' [gpus]   -- retries gpurun while the pod answers busy / transient
T=$1; CMD=$2; G=${3:-1}
for i in $(seq 1 12); do
  if [ "$G" = "1" ]; then /usr/local/graft/bin/gpurun --timeout $T -- "$CMD" > /tmp/gpurun_last.log 2>&1; rc=$?
  else /usr/local/graft/bin/gpurun --gpus $G --timeout $T -- "$CMD" > /tmp/gpurun_last.log 2>&1; rc=$?; fi
  if grep -q "status=transient\|nothing was charged" /tmp/gpurun_last.log || [ $rc -eq 3 ]; then
    echo "[gpu_retry] attempt $i: busy, sleeping 60s"; sleep 60; continue
  fi
  break
done
tail -60 /tmp/gpurun_last.log
exit $rc
