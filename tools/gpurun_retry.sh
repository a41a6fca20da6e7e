#!/bin/bash
# dev helper: retry a gpurun call while the pod answers "busy" (exit 3); usage: gpurun_retry.sh LOG [gpurun args...]
LOG=$1; shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun "$@" > "$LOG" 2>&1
  rc=$?
  if [ $rc -ne 3 ] && ! grep -q "status=transient" "$LOG"; then exit $rc; fi
  sleep 120
done
exit 3
