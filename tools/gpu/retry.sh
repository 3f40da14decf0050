#!/bin/bash
# usage: tools/gpu/retry.sh <timeout> <out-file> <command...>   -- retries gpurun while the pod answers busy (rc 3)
T=$1; OUT=$2; shift 2
for i in $(seq 1 20); do
  /usr/local/graft/bin/gpurun --timeout $T -- "$@" > $OUT 2>&1
  rc=$?
  if grep -q "status=transient" $OUT || [ $rc -eq 3 ]; then sleep 150; continue; fi
  break
done
tail -30 $OUT
