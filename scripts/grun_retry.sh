#!/bin/bash
# scripts/grun.sh with retries while the pod answers "transient" (busy): scripts/grun_retry.sh [--gpus N] <timeout> '<command>' <logfile>
cd "$(dirname "$0")/.."
LOG="${@: -1}"
ARGS=("${@:1:$#-1}")
for i in $(seq 1 40); do
  scripts/grun.sh "${ARGS[@]}" > "$LOG" 2>&1
  if ! grep -q "status=transient\|rc=None" "$LOG"; then break; fi
  sleep 120
done
tail -5 "$LOG"
