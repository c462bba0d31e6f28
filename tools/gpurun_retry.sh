#!/usr/bin/env bash
# gpurun with retries on "busy" (exit 3: nothing charged): tools/gpurun_retry.sh [gpurun args] -- '<command>'
for attempt in $(seq 1 40); do
  /usr/local/graft/bin/gpurun "$@"
  rc=$?
  if [ "$rc" != "3" ]; then exit $rc; fi
  echo "[retry] attempt $attempt answered busy; sleeping 90 s" >&2
  sleep 90
done
exit 3
