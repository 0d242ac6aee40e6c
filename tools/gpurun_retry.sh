#!/bin/bash
# tools/gpurun_retry.sh LOGFILE TIMEOUT 'command'  -- retry gpurun while the pod answers "busy" (exit 3)
log=$1; to=$2; shift 2
for attempt in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout $to -- "$@" > $log 2>&1
  rc=$?
  if ! grep -q "status=transient" $log; then exit $rc; fi
  sleep 90
done
exit 3
