#!/bin/bash
# gpurun with retries while the pod answers "busy" (exit 3).  Usage: tools/gpurun_retry.sh <timeout> <command...>
to=$1; shift
for n in 1 2 3 4 5 6 7 8 9 10; do
  /usr/local/graft/bin/gpurun --timeout $to -- "$@"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 150
done
exit 3
