#!/bin/bash
# one compute-sanitizer tool per gpurun call (B200_PROFILING.md): $1 = memcheck | racecheck
set -u
TOOL=${1:-memcheck}
O=gpurun_out
mkdir -p $O
ARG=""; [ "$TOOL" = "racecheck" ] && ARG="race"
if timeout 600 python tools/dev/dev_sanitize.py $ARG > $O/sanitize_plain_$TOOL.log 2>&1; then
  timeout 1500 compute-sanitizer --tool $TOOL --print-limit 20 python tools/dev/dev_sanitize.py $ARG > $O/sanitize_$TOOL.log 2>&1; echo "$TOOL rc=$?"
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|sanitize pass|Error|hazard" $O/sanitize_$TOOL.log | head -20
else
  echo "plain run failed"; tail -5 $O/sanitize_plain_$TOOL.log
fi
