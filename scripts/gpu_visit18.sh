#!/bin/bash
TAG=${1:-r1s}
OUT=gpurun_out/$TAG
mkdir -p $OUT
T0=$(date +%s); stamp() { echo "[t+$(( $(date +%s) - T0 ))s] $*"; }
stamp pytest; timeout 900 python -m pytest tests -q -m gpu --maxfail=20 > $OUT/pytest.log 2>&1; echo "pytest rc=$?"; grep -E "^(FAILED|ERROR)|passed|failed" $OUT/pytest.log | head -30
stamp small; timeout 600 python scripts/small_sizes.py > $OUT/small_sizes.txt 2>&1; echo "rc=$?"; grep -v "jit_\|launching\|Found CUDA" $OUT/small_sizes.txt | tail -40
stamp done
