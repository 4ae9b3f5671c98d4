#!/bin/bash
# usage: gpurun_retry.sh <tag> <timeout> <command...>  -- retries while the pod answers "transient"
tag=$1; to=$2; shift 2
for i in 1 2 3 4 5 6 7 8 9 10 11 12; do
  gpurun --timeout $to -- "$@" > gpurun_out/call_$tag.log 2>&1
  if grep -q "status=transient" gpurun_out/call_$tag.log; then sleep 45; continue; fi
  break
done
