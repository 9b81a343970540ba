#!/usr/bin/env bash
# Retry a gpurun call while the pod answers "busy" (exit 3).  Usage: tools/gpurun_retry.sh <logfile> <gpurun args...>
log="$1"; shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun "$@" > "$log" 2>&1
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 90
done
exit 3
