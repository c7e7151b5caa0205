#!/bin/bash
# usage: tools/gpurun_retry.sh <log> <timeout> [--gpus N] -- <command>   (retries while the pod answers busy/transient)
log=$1; to=$2; shift 2
for i in $(seq 1 30); do
    /usr/local/graft/bin/gpurun --timeout $to "$@" > $log 2>&1
    if grep -q "status=transient\|status=busy" $log; then sleep 90; continue; fi
    break
done
