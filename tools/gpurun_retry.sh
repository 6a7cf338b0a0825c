#!/bin/bash
# usage: tools/gpurun_retry.sh <timeout> <log> [--gpus N] : runs `bash tools/_call.sh` on a GPU box, retrying while the pod is busy
T=$1; LOG=$2; shift 2
for i in $(seq 1 20); do
  /usr/local/graft/bin/gpurun "$@" --timeout $T -- 'bash tools/_call.sh' > $LOG 2>&1
  if grep -q "status=transient\|nothing was charged" $LOG; then sleep 150; else break; fi
done
