#!/bin/bash
# usage: tools/gpurun_retry.sh <timeout> <gpus> '<command>'   - retries while the pod answers "transient" / busy
T=${1:-3000}; G=${2:-1}; shift 2
for i in $(seq 1 40); do
  if [ "$G" = "1" ]; then /usr/local/graft/bin/gpurun --timeout $T -- "$@" > /tmp/gpurun_last.log 2>&1; else /usr/local/graft/bin/gpurun --gpus $G --timeout $T -- "$@" > /tmp/gpurun_last.log 2>&1; fi
  rc=$?
  if grep -q "status=transient" /tmp/gpurun_last.log || [ $rc -eq 3 ]; then echo "try $i: transient, retrying in 45 s"; sleep 45; continue; fi
  break
done
tail -60 /tmp/gpurun_last.log
