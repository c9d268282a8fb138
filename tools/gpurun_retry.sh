#!/bin/bash
# usage: tools/gpurun_retry.sh <log> <timeout_s> <command...>   — retries while the pod answers "busy" (exit 3)
LOG=$1; shift; TO=$1; shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout $TO -- "$@" > $LOG 2>&1
  rc=$?
  echo "attempt $i exit $rc" >> $LOG.attempts
  if [ $rc -ne 3 ]; then echo "exit $rc" >> $LOG; exit $rc; fi
  sleep 90
done
echo "exit 3 (gave up)" >> $LOG
