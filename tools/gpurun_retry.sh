#!/bin/bash
# usage: tools/gpurun_retry.sh <timeout> '<command>'  -- retries while the pod answers busy/transient (nothing charged)
T=$1; shift
for i in 1 2 3 4 5 6 7 8; do
  OUT=$(/usr/local/graft/bin/gpurun --timeout $T "$@" 2>&1)
  echo "$OUT" | tail -40
  if echo "$OUT" | grep -q "status=transient\|status=busy\|retry in a few minutes\|no box"; then sleep 150; continue; fi
  break
done
