#!/bin/bash
# gpurun with retries while the pod answers "busy" (exit 3).  Usage: tools/gpurun_retry.sh <timeout> [--gpus N] <command...>
to=$1; shift
flags=""
if [ "$1" = "--gpus" ]; then flags="--gpus $2"; shift 2; fi
for n in 1 2 3 4 5 6 7 8 9 10 11 12; do
  /usr/local/graft/bin/gpurun $flags --timeout $to -- "$@"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 150
done
exit 3
