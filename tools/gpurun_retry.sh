#!/bin/bash
# gpurun with retries while the pod answers "busy / draining" (exit code 3: nothing charged). Usage: tools/gpurun_retry.sh <gpurun args...>
for attempt in 1 2 3 4 5 6 7 8; do
  /usr/local/graft/bin/gpurun "$@"; rc=$?
  [ $rc -ne 3 ] && exit $rc
  echo "[retry] attempt $attempt answered busy; sleeping 90 s" >&2; sleep 90
done
exit 3
