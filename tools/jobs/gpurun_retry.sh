#!/bin/bash
# usage: gpurun_retry.sh <out-file> <timeout> <command...>   -- retries while the pod answers busy (exit 3)
out=$1; shift; to=$1; shift
for i in $(seq 1 40); do
  gpurun --timeout $to -- "$@" > $out 2>&1
  rc=$?
  if grep -q "status=transient" $out; then sleep 120; continue; fi
  exit $rc
done
