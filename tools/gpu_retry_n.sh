#!/bin/bash
# usage: tools/gpu_retry_n.sh <gpus> <timeout> '<command>'
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --gpus "$1" --timeout "$2" -- "$3"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 100
done
exit 3
