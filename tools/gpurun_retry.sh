#!/bin/bash
# Retries a gpurun call while the pod answers "no box or slot free" (exit code 3): tools/gpurun_retry.sh [gpurun args...]
for attempt in $(seq 1 30); do
  /usr/local/graft/bin/gpurun "$@"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  echo "[retry] attempt $attempt: pod busy, sleeping 90 s" >&2
  sleep 90
done
exit 3
