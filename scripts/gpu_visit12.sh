#!/bin/bash
TAG=${1:-r1i}
OUT=gpurun_out/$TAG
mkdir -p $OUT
T0=$(date +%s); stamp() { echo "[t+$(( $(date +%s) - T0 ))s] $*"; }
stamp sweep-compress
for T in 128 3 253; do timeout 60 build/sweep_compress 30 $T > $OUT/sweep_compress_t$T.txt 2>&1; echo "rc=$?"; cat $OUT/sweep_compress_t$T.txt; done
stamp pytest; timeout 900 python -m pytest tests -q -m gpu -k "compress or smoke or launch" --maxfail=20 > $OUT/pytest.log 2>&1; echo "pytest rc=$?"; grep -E "^(FAILED|ERROR)|passed|failed" $OUT/pytest.log | head -30
stamp prims; timeout 300 python scripts/time_prims.py compress compress01 compress99 > $OUT/prims.log 2>&1; cat $OUT/prims.log
stamp done
