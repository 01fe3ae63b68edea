#!/bin/bash
# usage: tools/_gpurun_retry.sh <logfile> <gpurun args...>  — retries while the pod answers "busy" (exit 3, nothing charged)
log="$1"; shift
for i in $(seq 1 20); do
  /usr/local/graft/bin/gpurun "$@" > "$log" 2>&1
  rc=$?
  # 3 = pod busy; 2 with "already running" = one of our own calls is still in flight: wait for it
  if [ $rc -eq 2 ] && grep -q "already running" "$log"; then sleep 60; continue; fi
  if [ $rc -ne 3 ]; then echo "rc=$rc" >> "$log"; exit $rc; fi
  sleep 150
done
