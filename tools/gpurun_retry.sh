#!/bin/bash
# usage: tools/gpurun_retry.sh LOGFILE TIMEOUT -- command...   (retries while the pod answers "busy", exit code 3)
LOG=$1; TO=$2; shift 3
for i in $(seq 1 20); do
  /usr/local/graft/bin/gpurun --timeout $TO -- "$@" > $LOG 2>&1
  rc=$?
  [ $rc -ne 3 ] && exit $rc
  sleep 90
done
exit 3
