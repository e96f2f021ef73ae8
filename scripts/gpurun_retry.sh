#!/bin/bash
# gpurun with retries while the pod answers "busy / draining" (exit code 3, nothing charged).
# usage: scripts/gpurun_retry.sh LOGFILE [gpurun args...] -- 'command'
log="$1"; shift
for try in $(seq 1 40); do
  /usr/local/graft/bin/gpurun "$@" > "$log" 2>&1
  rc=$?
  if [ $rc -ne 3 ] && ! grep -q "status=transient" "$log"; then exit $rc; fi
  sleep 90
done
exit 3
