#!/usr/bin/env bash
# development aid: retry a gpurun call until the pod has a slot (exit code 3 / "transient" = nothing charged)
# usage: gpurun_retry.sh <log> <timeout-s> [--gpus N] -- '<command>'
LOG=$1; shift; TMO=$1; shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout "$TMO" "$@" > "$LOG" 2>&1
  if ! grep -q "status=transient\|status=busy\|rc=None" "$LOG"; then break; fi
  sleep 90
done
echo "[retry] done after $i attempt(s)" >> "$LOG"
