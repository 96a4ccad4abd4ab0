#!/bin/bash
# usage: scripts/dev/gpurun_retry.sh [gpurun flags ...] -- '<command>'
# Retries while the pod answers "busy" (exit code 3: nothing charged), every 90 s, up to 40 times.
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun "$@"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 90
done
exit 3
