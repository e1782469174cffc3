#!/bin/bash
# usage: tools/gpu/run_retry.sh <log file> <timeout seconds> <gpurun extra args...> -- <command>
# Retries a gpurun call while the pod answers "busy" (exit code 3 / transient), up to 12 times.
log=$1; shift; tmo=$1; shift
for i in $(seq 1 12); do
  /usr/local/graft/bin/gpurun --timeout $tmo "$@" > "$log" 2>&1
  rc=$?
  if grep -q "status=transient\|status=busy" "$log" || [ $rc -eq 3 ]; then sleep 90; continue; fi
  exit $rc
done
exit 3
