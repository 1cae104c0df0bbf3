#!/bin/bash
# usage: scripts/gpu_retry.sh [--gpus N] <timeout> <command...>   -- retries gpurun while it answers "busy" (exit 3)
extra=""
if [ "$1" = "--gpus" ]; then extra="--gpus $2"; shift 2; fi
t=$1; shift
for i in 1 2 3 4 5 6 7 8; do
  /usr/local/graft/bin/gpurun $extra --timeout "$t" -- "$@"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 90
done
exit 3
