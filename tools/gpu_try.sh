#!/bin/bash
# usage: [GPUS=N] tools/gpu_try.sh <logfile> <timeout_s> <command...>   -- retries gpurun while the pod answers "transient"/busy
log=$1; shift; to=$1; shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --gpus ${GPUS:-1} --timeout $to -- "$@" > "$log" 2>&1
  rc=$?
  if grep -q "status=transient\|rc=3\|no box\|retry in a few minutes" "$log" && ! grep -q "status=ok" "$log"; then sleep 90; continue; fi
  exit $rc
done
exit 3
