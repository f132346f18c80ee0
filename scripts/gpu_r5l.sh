#!/bin/bash
# Round-2 visit r5l (one GPU): compute-sanitizer over the r5 kernels, full-size incumbent parity test.
TAG=${1:-r5l}
OUT=gpurun_out/$TAG
mkdir -p $OUT
T0=$(date +%s); stamp() { echo "[t+$(( $(date +%s) - T0 ))s] $*"; }
for TOOL in memcheck racecheck; do
  stamp $TOOL; timeout 900 compute-sanitizer --tool $TOOL --print-limit 20 python scripts/sanitize_r5.py > $OUT/$TOOL.txt 2>&1
  echo "$TOOL rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|hazard|Invalid|sanitize_r5|Error|assert" $OUT/$TOOL.txt | head -12
done
stamp incumbent; timeout 900 python -m pytest tests/test_incumbent_gpu.py -q -m gpu > $OUT/pytest_incumbent.log 2>&1; echo "rc=$?"; tail -3 $OUT/pytest_incumbent.log
stamp done
