#!/bin/bash
# usage: [GPUS=n] gpurun_retry.sh <tag> <timeout> <command>  -- retries while the pod answers "transient"/busy
tag=$1; to=$2; shift 2
for i in $(seq 1 15); do
  if [ -n "$GPUS" ]; then
    gpurun --gpus $GPUS --timeout $to -- "$@" > gpurun_out/call_$tag.log 2>&1
  else
    gpurun --timeout $to -- "$@" > gpurun_out/call_$tag.log 2>&1
  fi
  if grep -q "status=transient\|rc=3\|busy" gpurun_out/call_$tag.log && ! grep -q "status=ok" gpurun_out/call_$tag.log; then sleep 45; continue; fi
  break
done
