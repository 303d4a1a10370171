#!/bin/bash
# One GPU call inside a hard time budget: steps given as "name|limit_seconds|command" lines on stdin, each run under
# `timeout`, its rc and duration logged to gpurun_out/<TAG>_steps.log.    usage: tools/gpu_call.sh BUDGET TAG < steps
BUDGET=${1:-1400}
TAG=${2:-call}
OUT=gpurun_out
mkdir -p $OUT
LOG=$OUT/${TAG}_steps.log
: > $LOG
while IFS='|' read -r name lim cmd; do
    [ -z "$name" ] && continue
    left=$(( BUDGET - SECONDS ))
    if [ "$left" -lt 10 ]; then echo "$name: skipped (only ${left}s left)" >> $LOG; continue; fi
    [ "$lim" -gt "$left" ] && lim=$left
    t0=$SECONDS
    timeout -k 5 "$lim" bash -c "$cmd"
    echo "$name: rc=$? in $(( SECONDS - t0 ))s (limit ${lim}s)" >> $LOG
done
cat $LOG
