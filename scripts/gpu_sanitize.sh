#!/bin/bash
# compute-sanitizer over one pass of every primitive (multi-tile sizes).
OUT=gpurun_out/${1:-r1_san}; mkdir -p $OUT
for TOOL in memcheck racecheck synccheck; do
  timeout 600 compute-sanitizer --tool $TOOL --print-limit 20 python scripts/sanitize_run.py > $OUT/$TOOL.txt 2>&1
  echo "$TOOL rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|hazard|Invalid|sanitize_run" $OUT/$TOOL.txt | head -12
done
