#!/bin/bash
# gpurun with retries while the pod is busy (exit code 3 = no slot, nothing charged).
#   tools/gpu_retry.sh [--gpus N] TIMEOUT 'command'     (log: stdout)
GP=""
if [ "$1" = "--gpus" ]; then GP="--gpus $2"; shift 2; fi
T=$1; shift
for attempt in $(seq 1 40); do
  /usr/local/graft/bin/gpurun $GP --timeout "$T" -- "python -m sgmcmcjax_b200._build > /dev/null 2>&1; $*"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  echo "[gpu_retry] attempt $attempt: busy, sleeping 45 s" >&2
  sleep 45
done
exit 3
