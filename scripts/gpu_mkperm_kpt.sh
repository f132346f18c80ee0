#!/bin/bash
# A/B of the unordered mkperm tile size (32 / 40 / 48 keys per thread; 60 = the experimental kernel with
# 16-bit staging entries) + its parity tests
TAG=${1:-r2o}
OUT=gpurun_out/$TAG
mkdir -p $OUT
T0=$(date +%s); stamp() { echo "[t+$(( $(date +%s) - T0 ))s] $*"; }
stamp "pytest KPT=48"; DRJIT_B200_MKPERM_KPT=48 timeout 150 python -m pytest tests -q -m gpu -k "mkperm" > $OUT/pytest_kpt48.log 2>&1; echo "pytest rc=$?"; grep -E "^(FAILED|ERROR)|passed|failed" $OUT/pytest_kpt48.log | head -10
stamp "pytest KPT=40"; DRJIT_B200_MKPERM_KPT=40 timeout 150 python -m pytest tests -q -m gpu -k "unordered_tiles or mkperm_baseline" > $OUT/pytest_kpt40.log 2>&1; echo "pytest rc=$?"; grep -E "^(FAILED|ERROR)|passed|failed" $OUT/pytest_kpt40.log | head -10
stamp "pytest KPT=60"; DRJIT_B200_MKPERM_KPT=60 timeout 150 python -m pytest tests -q -m gpu -k "mkperm" > $OUT/pytest_kpt60.log 2>&1; echo "pytest rc=$?"; grep -E "^(FAILED|ERROR)|passed|failed" $OUT/pytest_kpt60.log | head -10
for K in 32 40 48 60; do
  stamp "prims KPT=$K"; DRJIT_B200_MKPERM_KPT=$K timeout 60 python scripts/time_prims.py mkperm --reps 20 > $OUT/prims_kpt$K.log 2>&1; cat $OUT/prims_kpt$K.log
done
stamp done
