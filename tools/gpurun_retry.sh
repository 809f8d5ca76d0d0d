#!/bin/bash
# usage: gpurun_retry.sh <timeout-seconds> <log> <command...>   -- retries while the pod answers "transient" (nothing charged)
T=$1; L=$2; shift 2
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout $T -- "$@" > $L 2>&1
  grep -q "status=transient" $L || break
  sleep 150
done
