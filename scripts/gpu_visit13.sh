#!/bin/bash
TAG=${1:-r1k}
OUT=gpurun_out/$TAG
mkdir -p $OUT
T0=$(date +%s); stamp() { echo "[t+$(( $(date +%s) - T0 ))s] $*"; }
stamp pytest; timeout 900 python -m pytest tests -q -m gpu -k "prefix or smoke or launch" --maxfail=20 > $OUT/pytest.log 2>&1; echo "pytest rc=$?"; grep -E "^(FAILED|ERROR)|passed|failed" $OUT/pytest.log | head -30
stamp prims; timeout 300 python scripts/time_prims.py scan scan64 scanseg > $OUT/prims.log 2>&1; cat $OUT/prims.log
stamp done
